"""Per-call cost of small products: back-to-back calls (throughput) and single calls from an idle
stream (latency), this library vs cuBLAS through torch.matmul (float64)."""
import sys, time
import torch
sys.path.insert(0, ".")
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
for n in (64, 256, 512, 1024):
    A = torch.randn(n, n, dtype=torch.float64, device=dev); B = torch.randn(n, n, dtype=torch.float64, device=dev); C = torch.empty(n, n, dtype=torch.float64, device=dev)
    ours = lambda: pipe.dgemm(0, 0, n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)
    ref = lambda: torch.matmul(A, B, out=C)
    for name, fn in (("ours", ours), ("torch/cuBLAS", ref)):
        for _ in range(20): fn()
        torch.cuda.synchronize()
        reps = 2000
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        lat = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream); b.synchronize(); lat.append(a.elapsed_time(b) * 1e3)
        print(f"n={n:5d} {name:13s} host issue {t_issue/reps*1e6:6.2f} us/call, back-to-back {t_all/reps*1e6:6.2f} us/call "
              f"({2*n**3/(t_all/reps)/1e12:5.2f} TF), single-call event time best {min(lat):6.2f} us  kernel={pipe.last_kernel() if name=='ours' else ''}", flush=True)
