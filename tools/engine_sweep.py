"""Tensor-stream engine knobs against the measured host-link ceiling (1 GPU, or --gpus N in ONE
process through ec_tensor_stream_run_multi): chunk size (matrices per device chunk) x ring depth,
pinned host memory, 512^2 matrices.  Prints one line per setting and the link probe before/after.
    python tools/engine_sweep.py [--gpus N] [--count 2048]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import eigencuda_b200 as ec

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--count", type=int, default=2048, help="matrices per GPU")
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--chunks", default="4,8,16,32,64,128")
ap.add_argument("--rings", default="2,3,4")
args = ap.parse_args()

n, elems = args.n, args.n * args.n
devices = list(range(args.gpus))
total = args.count * args.gpus
pipe = ec.CudaPipeline(0)
print(json.dumps({"link_before": ec.host_link_probe(devices, 256 << 20, 5)}), flush=True)
pin_in, pin_out = ec.PinnedBuffer((elems, total)), ec.PinnedBuffer((elems, total))
pin_in.array[:] = np.random.default_rng(0).uniform(-1, 1, (elems, 1))
F = np.asfortranarray(np.random.default_rng(1).uniform(-1, 1, (n, n)))
ip = (C.c_void_p * total)(*[pin_in.ptr + 8 * elems * i for i in range(total)])
op = (C.c_void_p * total)(*[pin_out.ptr + 8 * elems * i for i in range(total)])


def call():
    if args.gpus == 1:
        pipe.tensor_stream_raw(ec.EC_STREAM_RIGHT, F.ctypes.data, n, n, ip, op, total, n, n)
    else:
        pipe.tensor_stream_raw_multi(devices, ec.EC_STREAM_RIGHT, F.ctypes.data, n, n, ip, op, total, n, n)


for ring in [int(x) for x in args.rings.split(",")]:
    for chunk in [int(x) for x in args.chunks.split(",")]:
        pipe.set_stream_config(chunk, ring, 0)
        call()
        best, tot, reps = 1e9, 0.0, 4
        for _ in range(reps):
            t0 = time.perf_counter()
            call()
            dt = time.perf_counter() - t0
            best, tot = min(best, dt), tot + dt
        print(json.dumps({"chunk": chunk or "default (64 MiB, ramped)", "ring": ring, "chunk_mib": chunk * elems * 8 / 2**20,
                          "matrices_per_s_avg": total * reps / tot, "matrices_per_s_best": total / best,
                          "gbs_each_way_avg": total * elems * 8 * reps / tot / 1e9}), flush=True)
print(json.dumps({"link_after": ec.host_link_probe(devices, 256 << 20, 5)}), flush=True)
# one direction at a time through plain copies of the engine's chunk size, for comparison
for mib in (8, 32, 128):
    print(json.dumps({"probe_mib": mib, **ec.host_link_probe(devices, mib << 20, 5)}), flush=True)
pipe.close()
