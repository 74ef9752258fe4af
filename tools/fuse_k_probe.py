"""Fused tile turn-around vs plain epilogue of the 128x128 tile as a function of K (short tiles): GB/s and TFLOP/s,
L2 flushed between calls.  Decides the dispatcher's threshold (fuse only from `EC_FUSE_MIN_KBLOCKS` k-iterations)."""
import os, sys, torch
sys.path.insert(0, ".")
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
os.environ["EC_FORCE_TILE"] = "1"
for fuse in ("1", "0"):
    os.environ["EC_FUSE_EPILOGUE"] = fuse
    os.environ["EC_FUSE_MIN_KBLOCKS"] = os.environ.get("PROBE_MIN_KBLOCKS", "1")   # 1: fuse at every K (the A/B); unset in the library: 2
    pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
    for (m, n) in ((1 << 20, 128), (8192, 8192)):
        for k in (16, 32, 48, 64, 96, 128):
            A = torch.randn(m * k, dtype=torch.float64, device=dev); B = torch.randn(k * n, dtype=torch.float64, device=dev)
            C = torch.empty(m * n, dtype=torch.float64, device=dev)
            fn = lambda: pipe.dgemm(0, 0, m, n, k, 1.0, A.data_ptr(), m, B.data_ptr(), k, 0.0, C.data_ptr(), m)
            fn(); torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); fn(); b.record(stream); b.synchronize(); best = min(best, a.elapsed_time(b))
            gb = 8.0 * (m * k + k * n + m * n) / 1e9
            print(f"fuse={fuse} {m}x{n}x{k}: {best*1e3:8.1f} us  {gb/best*1e3:7.0f} GB/s  {2.0*m*n*k/best/1e9:6.2f} TFLOP/s  {pipe.last_kernel()}", flush=True)
            del A, B, C
    pipe.close()
