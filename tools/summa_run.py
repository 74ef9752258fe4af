"""BASELINE.json configs[4]: one FP64 DGEMM, 2D block-partitioned over the GPUs of the box with
NCCL panel broadcasts (eigencuda_b200/summa.py).  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        --master-port 29533 tools/summa_run.py --size 65536 --panel 2048

Operands are generated on the device per owner block (34 GB per matrix at n = 65536).  Timing:
CUDA events around the whole SUMMA loop, barrier on both sides, max over ranks.  Verification:
sampled C elements against long-double dot products of generator rows/columns (CPU oracle).
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eigencuda_b200 as ec
from eigencuda_b200 import summa

SEED = 20200210
ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=16384, dest="n")
ap.add_argument("--panel", type=int, default=2048, dest="kb")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--samples", type=int, default=6)
ap.add_argument("--margin", type=int, default=0, help="SMs left free for NCCL")
ap.add_argument("--backend", default="native", choices=["native", "ozaki"], help="arithmetic of the local panel products")
ap.add_argument("--persistent", type=int, default=0, help="1: persistent grid + stream-K (the single-GPU default); 0: one CTA per tile, which coexists with NCCL kernels")
args = ap.parse_args()

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = args.n
L = summa.Layout(n, n, n, world, rank)
rg, cg = summa.make_groups(L) if world > 1 else (None, None)
stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(local, stream=stream.cuda_stream)
margin = args.margin
pipe.set_sm_margin(margin)
if not args.persistent:
    pipe.set_schedule(stream_k=False, one_cta_per_tile=True)
if args.backend == "ozaki":
    pipe.set_backend(ec.EC_BACKEND_OZAKI_I8)

(am, ak), (bk, bn), (cm, cn) = L.a_shape, L.b_shape, L.c_shape
A_loc = torch.empty((ak, am), dtype=torch.float64, device=dev)     # column-major am x ak
B_loc = torch.empty((bn, bk), dtype=torch.float64, device=dev)
C_loc = torch.empty((cn, cm), dtype=torch.float64, device=dev)
pipe.fill_uniform_block(SEED, 0, A_loc.data_ptr(), am, ak, am, L.m0, L.ka0, n)
pipe.fill_uniform_block(SEED, 1, B_loc.data_ptr(), bk, bn, bk, L.kb0, L.n0, n)


def local_gemm(mm, nn, kk, At, lda, Bt, ldb, beta, Ct, ldc):
    pipe.dgemm(0, 0, mm, nn, kk, 1.0, At.data_ptr(), lda, Bt.data_ptr(), ldb, beta, Ct.data_ptr(), ldc)


def run():
    summa.summa(L, A_loc, B_loc, C_loc, local_gemm, kb=args.kb, row_group=rg, col_group=cg)


run()                                   # warm-up (NCCL channels, kernel attributes)
torch.cuda.synchronize()
times = []
for _ in range(args.reps):
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run()
    e1.record(stream)
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
ms = min(times)

# sampled verification (outside the timed region)
from oracle import oracle_py as orc
rng = np.random.default_rng(1234 + rank)
worst = 0.0
for _ in range(args.samples):
    i, j = int(rng.integers(L.m0, L.m1)), int(rng.integers(L.n0, L.n1))
    arow = orc.uniform_strided(SEED, 0, i, n, n).astype(np.longdouble)        # A(i, :)
    bcol = orc.uniform_strided(SEED, 1, j * n, 1, n).astype(np.longdouble)    # B(:, j)
    want = float(np.dot(arow, bcol))
    got = float(C_loc[j - L.n0, i - L.m0].item())
    scale = float(np.linalg.norm(arow.astype(np.float64)) * np.linalg.norm(bcol.astype(np.float64)))
    worst = max(worst, abs(got - want) / scale)
w = torch.tensor([worst], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
if rank == 0:
    tf = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
    print(json.dumps({"workload": f"single DGEMM n={n}, 2D block partition {L.P}x{L.Q}, SUMMA panels kb={args.kb}, "
                                  "NCCL broadcasts double-buffered against the DMMA kernel",
                      "n_gpus": world, "ms": ms, "tflops": tf, "tflops_per_gpu": tf / world, "local_kernel": pipe.last_kernel(),
                      "frac_of_fp64_peak": tf / world / 37.22496,
                      "sampled_rel_err_vs_long_double": float(w.item()),
                      "backend": args.backend, "sm_margin": margin, "persistent": bool(args.persistent), "panel_bytes_received_per_gpu": (am * (n - ak) + (n - bk) * bn) * 8}), flush=True)
pipe.close()
if world > 1:
    dist.destroy_process_group()
