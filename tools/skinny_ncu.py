"""tools/skinny_ncu.py -- the HBM-bound skinny products for an ncu capture: 2^20 x 16 x 128 (128x16 tile kernel,
98 % of the HBM peak in bench.py's extras) and 2^20 x 128 x 16 (128x128 tile, one k-iteration per tile), two
launches each with the L2 flushed in between.  Run plainly first, then under
  ncu --set full --clock-control none -k regex:dgemm_tma -o gpurun_out/last/prof_skinny python tools/skinny_ncu.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (m, n, k) in ((1 << 20, 16, 128), (1 << 20, 128, 16)):
    A = torch.empty(m * k, dtype=torch.float64, device=dev)
    B = torch.empty(k * n, dtype=torch.float64, device=dev)
    C = torch.empty(m * n, dtype=torch.float64, device=dev)
    pipe.fill_uniform(1, 0, A.data_ptr(), m * k, 1); pipe.fill_uniform(1, 1, B.data_ptr(), k * n, 1)
    for _ in range(2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        pipe.dgemm(0, 0, m, n, k, 1.0, A.data_ptr(), m, B.data_ptr(), k, 0.0, C.data_ptr(), m)
        b.record(stream); b.synchronize()
        nbytes = 8.0 * (m * k + k * n + m * n)
        print(f"m={m} n={n} k={k} {pipe.last_kernel()} {a.elapsed_time(b):.4f} ms  {nbytes / a.elapsed_time(b) / 1e6:.0f} GB/s  algorithmic bytes {nbytes:.0f}", flush=True)
    del A, B, C
torch.cuda.synchronize(); print("done")
