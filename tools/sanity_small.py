"""Small pass over every kernel family, for compute-sanitizer (memcheck) runs:
    compute-sanitizer --tool memcheck python tools/sanity_small.py"""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import eigencuda_b200 as ec
from oracle import oracle_py as orc

pipe = ec.CudaPipeline(0)
def check(ta, tb, m, n, k, tag):
    A = orc.fill_uniform(3, 1, (k, m) if ta else (m, k)); B = orc.fill_uniform(3, 2, (n, k) if tb else (k, n))
    dA, dB, dC = ec.CudaMatrix(A, pipe.get_stream()), ec.CudaMatrix(B, pipe.get_stream()), ec.CudaMatrix(m, n, pipe.get_stream())
    pipe.dgemm(ta, tb, m, n, k, 1.0, dA.data(), A.shape[0], dB.data(), B.shape[0], 0.0, dC.data(), m)
    err = orc.rel_frobenius(dC.to_eigen(), orc.dgemm(A, B, ta, tb, flavour="fast"))
    print(tag, pipe.last_kernel(), (m, n, k), err, flush=True)
    assert err <= 1e-12
check(0, 0, 3, 2, 2, "simt")
check(1, 1, 65, 33, 17, "simt")
for tile in (1, 2, 3):
    os.environ["EC_FORCE_TILE"] = str(tile)
    for sk in (0, 2):
        os.environ["EC_STREAM_K"] = str(sk)
        check(0, 0, 200, 136, 330, f"tile{tile} sk{sk}")
        check(1, 0, 130, 258, 70, f"tile{tile} sk{sk}")
        check(0, 1, 384, 256, 160, f"tile{tile} sk{sk}")
os.environ.pop("EC_FORCE_TILE"); os.environ.pop("EC_STREAM_K")
# batched + tensor stream (pageable -> staged), every kind
F = orc.fill_uniform(3, 0, (96, 96)); ts = [orc.fill_uniform(3, i + 1, (96, 96)) for i in range(7)]
pipe.set_stream_config(2, 3, 2)
for kind, name in ((ec.EC_STREAM_RIGHT, "right"), (ec.EC_STREAM_LEFT_T, "left_t"), (ec.EC_STREAM_BASIS, "basis")):
    for r, w in zip(pipe.tensor_stream(kind, F, ts), orc.tensor_stream(name, F, ts, flavour="fast")):
        assert orc.rel_frobenius(r, w) <= 1e-12
    print("stream", name, "ok", flush=True)
# Ozaki backend
pipe.set_backend(ec.EC_BACKEND_OZAKI_I8)
check(0, 0, 256, 192, 320, "ozaki")
check(1, 1, 200, 136, 200, "ozaki")
pipe.set_backend(ec.EC_BACKEND_FP64_DMMA)
# generators
d = ec.CudaMatrix(100, 3, pipe.get_stream()); pipe.fill_uniform(1, 5, d.data(), 100, 3)
pipe.fill_uniform_block(1, 5, d.data(), 50, 3, 100, 7, 2, 300)
pipe.synchronize(); pipe.close()
print("sanity_small done")
