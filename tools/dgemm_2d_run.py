"""BASELINE.json configs[4] on its own: one FP64 DGEMM 2D-partitioned over the GPUs of the box
(ec_dgemm_2d: one process, NCCL panel broadcasts).  Same leg as bench.py's `dgemm_2d` block.
    python tools/dgemm_2d_run.py [--gpus N] [--size n] [--panel kb] [--backend native|ozaki]
                                 [--schedule persistent|tile]
Environment: EC_2D_NCCL_CTAS caps the CTAs NCCL's broadcast kernels may use (default 8, 0 = NCCL's own choice)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import eigencuda_b200 as ec
from oracle import oracle_py as orc

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=ec.count_available_gpus())
ap.add_argument("--size", type=int, default=0)
ap.add_argument("--panel", type=int, default=4096)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--backend", default="native", choices=["native", "ozaki"])
ap.add_argument("--schedule", default="default", choices=["default", "persistent", "tile"])
a = ap.parse_args()
sched = {"default": None, "persistent": (True, False), "tile": (False, True)}[a.schedule]
out = bench.dgemm_2d_leg(ec, orc, a.gpus, n=a.size or None, panel=a.panel, reps=a.reps,
                         backend=ec.EC_BACKEND_OZAKI_I8 if a.backend == "ozaki" else None, schedule=sched)
out.update({"backend": a.backend, "schedule": a.schedule, "nccl_ctas": os.environ.get("EC_2D_NCCL_CTAS", "8 (default)")})
print(json.dumps(out), flush=True)
