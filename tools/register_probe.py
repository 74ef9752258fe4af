"""tools/register_probe.py -- what does pinning caller memory in place cost (ec_host_register /
ec_host_unregister), per matrix size, from one thread and from several?  Decides whether the
tensor-stream engine could pin pageable tensors on the fly instead of staging them."""
import os, sys, time, threading
import ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigencuda_b200 as ec
lib = ec._lib.lib

pipe = ec.CudaPipeline()
for n in (512, 1000, 2048):
    count = 64
    arrs = [np.asfortranarray(np.random.default_rng(i).random((n, n))) for i in range(count)]
    for threads in (1, 4, 8):
        def work(lo, hi, out):
            t0 = time.perf_counter()
            for a in arrs[lo:hi]:
                lib.ec_host_register(C.c_void_p(a.ctypes.data), a.nbytes)
            t1 = time.perf_counter()
            for a in arrs[lo:hi]:
                lib.ec_host_unregister(C.c_void_p(a.ctypes.data))
            t2 = time.perf_counter()
            out.append((t1 - t0, t2 - t1))
        res = []
        ths = [threading.Thread(target=work, args=(count * t // threads, count * (t + 1) // threads, res)) for t in range(threads)]
        t0 = time.perf_counter()
        for t in ths: t.start()
        for t in ths: t.join()
        wall = time.perf_counter() - t0
        reg = max(r[0] for r in res); unreg = max(r[1] for r in res)
        print(f"n={n} ({n*n*8/2**20:.1f} MiB) threads={threads}: register {1e6*reg/ (count/threads):7.1f} us/call/thread, unregister {1e6*unreg/(count/threads):7.1f}; "
              f"wall per matrix (reg+unreg) {1e6*wall/count:7.1f} us  -> {n*n*8/ (wall/count)/1e9:6.1f} GB/s equivalent", flush=True)
