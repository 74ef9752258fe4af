import os, sys, torch
sys.path.insert(0, ".")
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
for n in (128, 192, 256, 384, 512, 640, 768, 896, 1024, 1280, 1536, 1792, 2048, 2560, 3072, 4096):
    A = torch.randn(n*n, dtype=torch.float64, device=dev); B = torch.randn(n*n, dtype=torch.float64, device=dev); C = torch.empty(n*n, dtype=torch.float64, device=dev)
    row = []
    for tile, sk in ((0, 1), (2, 0), (2, 2), (3, 0), (3, 2), (4, 0), (4, 2), (1, 0), (1, 2)):
        if tile: os.environ["EC_FORCE_TILE"] = str(tile)
        else: os.environ.pop("EC_FORCE_TILE", None)
        os.environ["EC_STREAM_K"] = str(sk)
        fn = lambda: pipe.dgemm(0,0,n,n,n,1.0,A.data_ptr(),n,B.data_ptr(),n,0.0,C.data_ptr(),n)
        for _ in range(5): fn()
        torch.cuda.synchronize()
        # back-to-back throughput (kernel time without the launch latency of an idle stream)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(50): fn()
        b.record(stream); b.synchronize()
        ms = a.elapsed_time(b) / 50
        row.append(f"{'auto' if not tile else 'LSXT'[tile-1]}{'sk' if sk==2 else ''}:{2*n**3/ms/1e9:5.1f}({pipe.last_kernel().replace('tma_dmma_NN','') or 'L'})")
    print(n, " ".join(row), flush=True)
