"""Debug probe: full-output comparison of every tile config x schedule against L/data-parallel."""
import os, sys
import torch
sys.path.insert(0, ".")
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
n, batch = int(sys.argv[1]), int(sys.argv[2])
A = torch.empty(n*n*batch, dtype=torch.float64, device=dev); B = torch.empty(n*n, dtype=torch.float64, device=dev)
pipe.fill_uniform(1, 1, A.data_ptr(), n*n, batch); pipe.fill_uniform(1, 0, B.data_ptr(), n*n, 1)
def run(tile, sk):
    os.environ["EC_FORCE_TILE"] = str(tile); os.environ["EC_STREAM_K"] = str(sk)
    Cm = torch.full((n*n*batch,), float("nan"), dtype=torch.float64, device=dev)
    pipe.dgemm_strided_batched(0,0,n,n,n,1.0,A.data_ptr(),n,n*n,B.data_ptr(),n,0,0.0,Cm.data_ptr(),n,n*n,batch)
    torch.cuda.synchronize()
    return Cm, pipe.last_kernel()
ref, _ = run(1, 0)
for tile in (1, 2, 3):
    for sk in (0, 2):
        for rep in range(6):
            Cm, name = run(tile, sk)
            diff = (Cm - ref).abs()
            bad = torch.nonzero(~(diff < 1e-9)).flatten()
            msg = ""
            if bad.numel():
                idx = bad[:5].tolist()
                msg = " first bad: " + str([(i // (n*n), (i % (n*n)) % n, (i % (n*n)) // n, float(Cm[i]), float(ref[i])) for i in idx]) + f" nbad={bad.numel()} mats={sorted(set((bad // (n*n)).tolist()))[:12]}"
            print(tile, sk, rep, name, "maxdiff", float(diff.nan_to_num(1e30).max()), msg[:160], flush=True)
