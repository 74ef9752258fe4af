"""First contact for EC_BACKEND_OZAKI_I8: accuracy vs the long-double oracle and speed vs native."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
import eigencuda_b200 as ec
from oracle import oracle_py as orc
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
stage = sys.argv[1] if len(sys.argv) > 1 else "acc"
if stage == "acc":
    for (m, n, k, ta, tb) in [(128, 64, 64, 0, 0), (128, 64, 128, 0, 0), (256, 192, 320, 0, 0), (200, 136, 520, 0, 0), (256, 128, 256, 1, 0), (256, 128, 256, 0, 1), (384, 200, 1000, 1, 1), (512, 512, 512, 0, 0)]:
        A = orc.fill_uniform(7, 1, (k, m) if ta else (m, k)); B = orc.fill_uniform(7, 2, (n, k) if tb else (k, n))
        dA, dB, dC = ec.CudaMatrix(A, pipe.get_stream()), ec.CudaMatrix(B, pipe.get_stream()), ec.CudaMatrix(m, n, pipe.get_stream())
        want = orc.dgemm(A, B, ta, tb, flavour="ld")
        for backend in (ec.EC_BACKEND_OZAKI_I8, ec.EC_BACKEND_FP64_DMMA):
            pipe.set_backend(backend)
            print("launch", m, n, k, ta, tb, backend, flush=True)
            pipe.dgemm(ta, tb, m, n, k, 1.0, dA.data(), A.shape[0], dB.data(), B.shape[0], 0.0, dC.data(), m)
            pipe.synchronize()
            got = dC.to_eigen()
            print("  ", pipe.last_kernel(), "rel frob err", orc.rel_frobenius(got, want), flush=True)
else:
    n, batch = 512, 2048
    A = torch.empty(n*n*batch, dtype=torch.float64, device=dev); B = torch.empty(n*n, dtype=torch.float64, device=dev); Cm = torch.empty(n*n*batch, dtype=torch.float64, device=dev)
    pipe.fill_uniform(1, 1, A.data_ptr(), n*n, batch); pipe.fill_uniform(1, 0, B.data_ptr(), n*n, 1)
    def timeit(fn, reps=4):
        for _ in range(2): fn()
        torch.cuda.synchronize(); best = 1e9
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream); b.synchronize(); best = min(best, a.elapsed_time(b))
        return best
    for backend, name in ((ec.EC_BACKEND_FP64_DMMA, "native"), (ec.EC_BACKEND_OZAKI_I8, "ozaki")):
        pipe.set_backend(backend)
        ms = timeit(lambda: pipe.dgemm_strided_batched(0,0,n,n,n,1.0,A.data_ptr(),n,n*n,B.data_ptr(),n,0,0.0,Cm.data_ptr(),n,n*n,batch))
        print(name, pipe.last_kernel(), f"batched 512^3 x {batch}: {ms:.3f} ms, {batch/ms*1e3:.0f} matrices/s, {2*n**3*batch/ms/1e9:.1f} TF-equivalent", flush=True)
    for nn in (2048, 4096, 8192):
        A2 = torch.empty(nn*nn, dtype=torch.float64, device=dev); B2 = torch.empty(nn*nn, dtype=torch.float64, device=dev); C2 = torch.empty(nn*nn, dtype=torch.float64, device=dev)
        pipe.fill_uniform(1, 1, A2.data_ptr(), nn*nn, 1); pipe.fill_uniform(1, 0, B2.data_ptr(), nn*nn, 1)
        for backend, name in ((ec.EC_BACKEND_FP64_DMMA, "native"), (ec.EC_BACKEND_OZAKI_I8, "ozaki")):
            pipe.set_backend(backend)
            ms = timeit(lambda: pipe.dgemm(0,0,nn,nn,nn,1.0,A2.data_ptr(),nn,B2.data_ptr(),nn,0.0,C2.data_ptr(),nn), 3)
            print(name, f"n={nn}: {ms:.3f} ms, {2*nn**3/ms/1e9:.1f} TF-equivalent", flush=True)
        del A2, B2, C2
