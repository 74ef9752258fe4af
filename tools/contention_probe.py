"""A product while another kernel holds some of the SMs (ec_debug_hold_sms on a side stream):
how long it takes with the static and with the dynamic schedules, and what the dispatcher notices.
    python tools/contention_probe.py ; EC_DYNAMIC_TILES=0 python tools/contention_probe.py"""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import eigencuda_b200 as ec
side = torch.cuda.Stream()
t0 = time.perf_counter(); ec.debug_hold_sms(0, side.cuda_stream, 24, 30000); side.synchronize()
print("hold kernel alone: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
for n in (4096, 2048, 1024):
    p = ec.CudaPipeline(0)
    A = torch.randn(n * n, dtype=torch.float64, device="cuda"); B = torch.randn(n * n, dtype=torch.float64, device="cuda"); C = torch.empty(n * n, dtype=torch.float64, device="cuda")
    run = lambda: p.dgemm(0, 0, n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)
    for _ in range(3): run()
    p.synchronize()
    t0 = time.perf_counter(); run(); p.synchronize(); quiet = (time.perf_counter() - t0) * 1e3
    row = []
    for rep in range(4):
        ec.debug_hold_sms(0, side.cuda_stream, 24, 40000)
        time.sleep(0.003)
        t0 = time.perf_counter(); run(); p.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        side.synchronize()
        row.append("%.2f ms (%s, events %d)" % (dt, p.last_kernel(), p.contention_events()))
    print(f"n={n} EC_DYNAMIC_TILES={os.environ.get('EC_DYNAMIC_TILES','1')} quiet {quiet:.2f} ms | under a 40 ms hold of 24 SMs: " + " | ".join(row), flush=True)
    p.close()
