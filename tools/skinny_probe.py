"""tools/skinny_probe.py -- skinny / small-k shapes, where the product is HBM-bound rather than
DMMA-bound: achieved HBM GB/s (algorithmic bytes: read A and B once, write C once) next to TFLOP/s,
for the native path under each tile configuration, cuBLAS (torch.matmul) beside it.
Run on a B200: python tools/skinny_probe.py"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6458.4
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()                                  # operands larger than L2 anyway; keep it honest
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps

# (m, n, k): C(m x n) = A(m x k) * B(k x n), column-major
shapes = [(1 << 20, 32, 32), (1 << 20, 64, 64), (1 << 19, 128, 128), (1 << 20, 16, 128), (1 << 20, 16, 64), (1 << 21, 8, 128), (1 << 20, 128, 16),
          (64, 64, 1 << 20), (128, 128, 1 << 19), (8192, 8192, 64), (8192, 8192, 128)]
for (m, n, k) in shapes:
    A = torch.randn(m * k, dtype=torch.float64, device=dev)
    B = torch.randn(k * n, dtype=torch.float64, device=dev)
    C = torch.empty(m * n, dtype=torch.float64, device=dev)
    bytes_ = 8.0 * (m * k + k * n + m * n); flop = 2.0 * m * n * k
    row = []
    for tile in (0, 1, 2, 3, 4, 5):
        if tile: os.environ["EC_FORCE_TILE"] = str(tile)
        else: os.environ.pop("EC_FORCE_TILE", None)
        ms = timed(lambda: pipe.dgemm(0, 0, m, n, k, 1.0, A.data_ptr(), m, B.data_ptr(), k, 0.0, C.data_ptr(), m))
        row.append(f"{'auto' if not tile else 'LSXTN'[tile - 1]}: {bytes_ / ms / 1e6:6.0f} GB/s {flop / ms / 1e9:5.1f} TF ({pipe.last_kernel().replace('tma_dmma_NN', '') or 'L'})")
    os.environ.pop("EC_FORCE_TILE", None)
    # cuBLAS through torch: row-major views of the same column-major data: C^T = B^T A^T
    At = A.view(k, m); Bt = B.view(n, k); Ct = C.view(n, m)
    ms = timed(lambda: torch.matmul(Bt, At, out=Ct))
    row.append(f"cuBLAS: {bytes_ / ms / 1e6:6.0f} GB/s {flop / ms / 1e9:5.1f} TF")
    print(f"m={m} n={n} k={k} AI={flop / bytes_:5.1f} flop/B (HBM peak {HBM:.0f} GB/s): " + " | ".join(row), flush=True)
    del A, B, C
