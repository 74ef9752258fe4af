"""Staged first contact with the GPU: each stage prints before it launches, so that a hang or a
fault is attributable; run each stage under its own `timeout`."""
import sys
import numpy as np
sys.path.insert(0, ".")
import eigencuda_b200 as ec
from oracle import oracle_py as orc

stage = sys.argv[1]
pipe = ec.CudaPipeline(0)
print("stage", stage, flush=True)
if stage == "simt":
    A = np.asfortranarray(np.arange(1., 7.).reshape(3, 2)); B = np.asfortranarray(np.array([[1., 2.], [3., 4.]]))
    dA, dB, dC = ec.CudaMatrix(A, pipe.get_stream()), ec.CudaMatrix(B, pipe.get_stream()), ec.CudaMatrix(3, 2, pipe.get_stream())
    pipe.gemm(dA, dB, dC)
    print(pipe.last_kernel(), dC.to_eigen().tolist(), (A @ B).tolist(), flush=True)
elif stage.startswith("tma"):
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    for ta in (0, 1):
        for tb in (0, 1):
            A = orc.fill_uniform(1, 2, (n, n)); B = orc.fill_uniform(1, 3, (n, n))
            dA, dB, dC = ec.CudaMatrix(A, pipe.get_stream()), ec.CudaMatrix(B, pipe.get_stream()), ec.CudaMatrix(n, n, pipe.get_stream())
            print("launch", ta, tb, flush=True)
            pipe.dgemm(ta, tb, n, n, n, 1.0, dA.data(), n, dB.data(), n, 0.0, dC.data(), n)
            pipe.synchronize()
            got = dC.to_eigen()
            want = (A.T if ta else A) @ (B.T if tb else B)
            print(pipe.last_kernel(), "rel err", orc.rel_frobenius(got, want), flush=True)
pipe.close()
print("done", stage, flush=True)
