"""tools/compat_probe.py -- phase timing of the reference's README loop (copy_to_gpu -> gemm ->
conversion back, pageable memory) through this repo's compat API, with and without the staging
workers (EC_COMPAT_THREADS).  Run on a B200: python tools/compat_probe.py [n] [count]"""
import os, resource, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigencuda_b200 as ec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
count = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rng = np.random.default_rng(1)
F = np.asfortranarray(rng.random((n, n)))
ts = [np.asfortranarray(rng.random((n, n))) for _ in range(count)]

for threads in ("1", "4", None, None, None):
    if threads is None:
        os.environ.pop("EC_COMPAT_THREADS", None)
    else:
        os.environ["EC_COMPAT_THREADS"] = threads
    pipe = ec.CudaPipeline()
    A = ec.CudaMatrix(F, pipe.get_stream())
    B = ec.CudaMatrix(n, n, pipe.get_stream())
    Cm = ec.CudaMatrix(n, n, pipe.get_stream())
    for t in ts[:8]:
        B.copy_to_gpu(t); pipe.gemm(B, A, Cm); Cm.to_eigen()
    up = mm = dn = al = 0.0
    flt0 = resource.getrusage(resource.RUSAGE_SELF).ru_minflt
    t_all = time.perf_counter()
    for t in ts:
        t0 = time.perf_counter(); B.copy_to_gpu(t)
        t1 = time.perf_counter(); pipe.gemm(B, A, Cm)
        t2 = time.perf_counter(); r = np.empty((n, n), order="F")
        t2b = time.perf_counter(); ec._lib.check(ec._lib.lib.ec_matrix_download(Cm._h, r.ctypes.data))
        t3 = time.perf_counter()
        up += t1 - t0; mm += t2 - t1; dn += t3 - t2; al += t2b - t2
    dt = time.perf_counter() - t_all
    flt = (resource.getrusage(resource.RUSAGE_SELF).ru_minflt - flt0) / count
    ok = np.allclose(r, ts[-1] @ F)
    print(f"threads={threads or 'default'} n={n}: {count / dt:8.0f} matrices/s  per-iteration us: upload {1e6 * up / count:6.1f}"
          f" gemm {1e6 * mm / count:6.1f} convert {1e6 * dn / count:6.1f} (np.empty {1e6 * al / count:5.1f})  minor faults/iter {flt:6.1f}  ok={ok}", flush=True)
    del A, B, Cm, pipe

# callers that construct their matrices inside the loop: allocation cost with and without the pool
for pool in ("0", "1"):
    os.environ["EC_MEM_POOL"] = pool
    pipe = ec.CudaPipeline()
    A = ec.CudaMatrix(F, pipe.get_stream())
    for rep in range(2):
        t0 = time.perf_counter()
        for t in ts:
            B = ec.CudaMatrix(t, pipe.get_stream())
            Cm = ec.CudaMatrix(n, n, pipe.get_stream())
            pipe.gemm(B, A, Cm)
            r = Cm.to_eigen()
            B.close(); Cm.close()
        dt = time.perf_counter() - t0
    print(f"matrices constructed per iteration, EC_MEM_POOL={pool}: {count / dt:8.0f} matrices/s  ok={np.allclose(r, ts[-1] @ F)}", flush=True)
    del A, pipe
os.environ.pop("EC_MEM_POOL", None)
