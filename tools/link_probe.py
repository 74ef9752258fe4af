"""Host-link probe: pinned H2D + D2H running at once on every visible GPU simultaneously
(one thread per GPU), to see whether GPUs share a PCIe uplink / host memory bandwidth."""
import sys, threading, time
import torch
n = torch.cuda.device_count()
gpus = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(range(n))
SZ = 256 << 20
res = {}
def work(g, barrier):
    torch.cuda.set_device(g)
    hin = torch.empty(SZ, dtype=torch.uint8).pin_memory(); hout = torch.empty(SZ, dtype=torch.uint8).pin_memory()
    din = torch.empty(SZ, dtype=torch.uint8, device=f"cuda:{g}"); dout = torch.empty(SZ, dtype=torch.uint8, device=f"cuda:{g}")
    s1, s2 = torch.cuda.Stream(g), torch.cuda.Stream(g)
    torch.cuda.synchronize(g)
    barrier.wait()
    t0 = time.perf_counter()
    reps = 8
    for _ in range(reps):
        with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
    s1.synchronize(); s2.synchronize()
    dt = time.perf_counter() - t0
    res[g] = SZ * reps / dt / 1e9
b = threading.Barrier(len(gpus))
ts = [threading.Thread(target=work, args=(g, b)) for g in gpus]
[t.start() for t in ts]; [t.join() for t in ts]
print("gpus", gpus, "GB/s each way per GPU:", {g: round(v, 1) for g, v in sorted(res.items())}, "sum", round(sum(res.values()), 1))
