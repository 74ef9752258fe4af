"""A/B of one shape under a forced tile / schedule: back-to-back average in microseconds.
    EIGENCUDA_B200_LIB=... EC_DYNAMIC_TILES=.. EC_PDL=.. python tools/ab_probe.py"""
import os, sys, torch
sys.path.insert(0, ".")
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
out = []
for (n, tile, sk) in [tuple(int(v) for v in x.split(':')) for x in os.environ.get('AB_CASES', '1024:2:0,1024:2:2,512:4:0,640:4:0,2048:1:2,2048:1:0,4096:1:0').split(',')]:
    os.environ["EC_FORCE_TILE"] = str(tile); os.environ["EC_STREAM_K"] = str(sk)
    A = torch.randn((n, n), dtype=torch.float64, device=dev); B = torch.randn((n, n), dtype=torch.float64, device=dev)
    C = torch.empty((n, n), dtype=torch.float64, device=dev)
    fn = lambda: pipe.dgemm(0, 0, n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)
    for _ in range(5): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(50): fn()
        b.record(stream); b.synchronize()
        best = min(best, a.elapsed_time(b) / 50 * 1e3)
    one = 1e9
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); b.synchronize()
        one = min(one, a.elapsed_time(b) * 1e3)
    out.append(f"n={n} tile={tile} sk={sk}: {best:7.1f}/{one:6.1f} ({pipe.last_kernel()})")
print(f"lib={'old' if os.environ.get('EIGENCUDA_B200_LIB') else 'new'} DYN={os.environ.get('EC_DYNAMIC_TILES','1')} PDL={os.environ.get('EC_PDL','1')} | " + " | ".join(out), flush=True)
