"""Small squares, back to back and one at a time, against cuBLAS (torch.mm float64), per tile /
schedule choice.  EC_PDL=0 turns programmatic dependent launch off for comparison.
    python tools/small_probe2.py"""
import os, sys, torch
sys.path.insert(0, ".")
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)


def timed(fn, reps):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): fn()
    b.record(stream); b.synchronize()
    b2b = a.elapsed_time(b) / reps
    best = 1e9
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return b2b * 1e3, best * 1e3   # microseconds


for n in (128, 256, 384, 512, 640, 768, 1024, 2048):
    A = torch.randn((n, n), dtype=torch.float64, device=dev); B = torch.randn((n, n), dtype=torch.float64, device=dev)
    C = torch.empty((n, n), dtype=torch.float64, device=dev); Cb = torch.empty((n, n), dtype=torch.float64, device=dev)
    row = []
    cb2b, cone = timed(lambda: torch.mm(B, A, out=Cb), 50)
    row.append(f"cublas b2b {cb2b:6.1f} one {cone:6.1f} us |")
    for tile, sk in ((0, 1), (2, 0), (2, 2), (3, 0), (4, 0), (1, 2)):
        if tile: os.environ["EC_FORCE_TILE"] = str(tile)
        else: os.environ.pop("EC_FORCE_TILE", None)
        os.environ["EC_STREAM_K"] = str(sk)
        fn = lambda: pipe.dgemm(0, 0, n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)
        b2b, one = timed(fn, 50)
        name = ('auto' if not tile else 'LSXT'[tile - 1]) + ('sk' if sk == 2 else '')
        row.append(f"{name}({pipe.last_kernel().replace('tma_dmma_NN', '') or 'L'}) {b2b:6.1f}/{one:6.1f}")
    print(f"n={n:5d} PDL={os.environ.get('EC_PDL', '1')} ", "  ".join(row), flush=True)
