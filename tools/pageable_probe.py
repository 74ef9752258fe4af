"""e2e from PAGEABLE host memory (what std::vector<Eigen::MatrixXd> is): staging threads / chunk sweep."""
import ctypes as C, sys, time, os
import numpy as np
sys.path.insert(0, ".")
import eigencuda_b200 as ec
from oracle import oracle_py as orc
pipe = ec.CudaPipeline(0)
n, count = 512, 1024
elems = n * n
F = orc.fill_uniform(1, 0, (n, n))
base = np.empty((elems, count), dtype=np.float64, order="F"); base[:] = 0.5
outs = np.empty((elems, count), dtype=np.float64, order="F")
ip = (C.c_void_p * count)(*[base.ctypes.data + 8 * elems * i for i in range(count)])
op = (C.c_void_p * count)(*[outs.ctypes.data + 8 * elems * i for i in range(count)])
print("cpus", len(os.sched_getaffinity(0)))
for threads in (4, 8, 16, 24, 32):
    for chunk in (16, 64):
        pipe.set_stream_config(chunk, 3, threads)
        pipe.tensor_stream_raw(ec.EC_STREAM_RIGHT, F.ctypes.data, n, n, ip, op, count, n, n)
        t0 = time.perf_counter()
        for _ in range(3):
            pipe.tensor_stream_raw(ec.EC_STREAM_RIGHT, F.ctypes.data, n, n, ip, op, count, n, n)
        dt = (time.perf_counter() - t0) / 3
        print(f"threads {threads:2d} chunk {chunk:3d}: {count/dt:8.0f} matrices/s, {count*elems*8/dt/1e9:5.1f} GB/s each way", flush=True)
