"""tools/batched_small_probe.py -- the tensor-stream shape (a batch of n x n matrices against one
fixed n x n matrix, resident in HBM) for SMALL n, where a matrix is a tile or less: TFLOP/s and
achieved HBM GB/s (read T_i, write R_i) next to cuBLAS strided-batched (torch.matmul broadcasting).
Run on a B200: python tools/batched_small_probe.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

for n in (16, 32, 48, 64, 96, 128, 192, 256, 384):
    batch = max(64, (1 << 27) // (n * n))          # 1 GiB of tensor
    T = torch.randn(batch * n * n, dtype=torch.float64, device=dev)
    F = torch.randn(n * n, dtype=torch.float64, device=dev)
    R = torch.empty(batch * n * n, dtype=torch.float64, device=dev)
    flop = 2.0 * n ** 3 * batch; nbytes = 16.0 * n * n * batch
    row = []
    for tile in (0, 1, 2, 3, 4):
        if tile: os.environ["EC_FORCE_TILE"] = str(tile)
        else: os.environ.pop("EC_FORCE_TILE", None)
        ms = timed(lambda: pipe.dgemm_strided_batched(0, 0, n, n, n, 1.0, T.data_ptr(), n, n * n, F.data_ptr(), n, 0, 0.0,
                                                      R.data_ptr(), n, n * n, batch))
        row.append(f"{'auto' if not tile else 'LSXT'[tile - 1]}: {flop / ms / 1e9:5.1f} TF {nbytes / ms / 1e6:5.0f} GB/s ({pipe.last_kernel().replace('tma_dmma_NN', '') or 'L'})")
    os.environ.pop("EC_FORCE_TILE", None)
    # cuBLAS: column-major R_i = T_i F  <=>  row-major views R_i^T = F^T T_i^T
    Tt = T.view(batch, n, n); Ft = F.view(n, n); Rt = R.view(batch, n, n)
    ms = timed(lambda: torch.matmul(Ft, Tt, out=Rt))
    row.append(f"cuBLAS: {flop / ms / 1e9:5.1f} TF {nbytes / ms / 1e6:5.0f} GB/s")
    print(f"n={n} batch={batch}: " + " | ".join(row), flush=True)
    del T, R, F
