"""Stream-K vs whole-tile schedule of the 128x128 tile around n = 3072 (where the cost model's choice lost in
tools/small_probe.py), with the dynamic claims / fused turn-around switched on and off."""
import os, sys, torch
sys.path.insert(0, ".")
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2560, 2816, 3072, 3328, 3584, 4096, 5120, 6144]
ENVS = ({}, {"EC_DYNAMIC_TILES": "0"}, {"EC_FUSE_EPILOGUE": "0"}) if os.environ.get("SK_PROBE_ALL") else ({}, {"EC_DYNAMIC_TILES": "0"})
for env in ENVS:
    for k, v in env.items(): os.environ[k] = v
    pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
    for n in sizes:
        A = torch.randn(n*n, dtype=torch.float64, device=dev); B = torch.randn(n*n, dtype=torch.float64, device=dev); C = torch.empty(n*n, dtype=torch.float64, device=dev)
        row = []
        for tile, sk in ((0, 1), (1, 0), (1, 2), (2, 2)):
            if tile: os.environ["EC_FORCE_TILE"] = str(tile)
            else: os.environ.pop("EC_FORCE_TILE", None)
            os.environ["EC_STREAM_K"] = str(sk)
            fn = lambda: pipe.dgemm(0,0,n,n,n,1.0,A.data_ptr(),n,B.data_ptr(),n,0.0,C.data_ptr(),n)
            for _ in range(3): fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(10): fn()
            b.record(stream); b.synchronize()
            ms = a.elapsed_time(b) / 10
            row.append(f"{'auto' if not tile else 'LSXT'[tile-1]}{'sk' if sk==2 else ''}:{2*n**3/ms/1e9:5.1f}({pipe.last_kernel().replace('tma_dmma_NN','') or 'L'})")
        print(env, n, " ".join(row), flush=True)
        del A, B, C
    pipe.close()
    for k in env: os.environ.pop(k)
os.environ.pop("EC_FORCE_TILE", None); os.environ.pop("EC_STREAM_K", None)
