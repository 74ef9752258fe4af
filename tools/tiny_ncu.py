"""tools/tiny_ncu.py -- the HBM-bound end of the path for an ncu capture: a tensor of 131072
matrices of 32 x 32 against a fixed one (32x32 tile kernel) and the skinny 2^20 x 32 x 32 product
(64x32 tile kernel), two launches each.  Run plainly first, then under
  ncu --set full --clock-control none -k regex:dgemm_tma -o gpurun_out/prof_tiny python tools/tiny_ncu.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
n, batch = 32, 131072
T = torch.empty(batch * n * n, dtype=torch.float64, device=dev)
F = torch.empty(n * n, dtype=torch.float64, device=dev)
R = torch.empty(batch * n * n, dtype=torch.float64, device=dev)
pipe.fill_uniform(1, 1, T.data_ptr(), n * n, batch); pipe.fill_uniform(1, 0, F.data_ptr(), n * n, 1)
for _ in range(2):
    pipe.dgemm_strided_batched(0, 0, n, n, n, 1.0, T.data_ptr(), n, n * n, F.data_ptr(), n, 0, 0.0, R.data_ptr(), n, n * n, batch)
print(pipe.last_kernel())
m = 1 << 20
for _ in range(2):
    pipe.dgemm(0, 0, m, n, n, 1.0, T.data_ptr(), m, F.data_ptr(), n, 0.0, R.data_ptr(), m)
print(pipe.last_kernel())
torch.cuda.synchronize(); print("done")
