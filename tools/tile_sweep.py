"""Times every CTA tile configuration of dgemm_tma_dmma_kernel on square and batched problems
(EC_FORCE_TILE=1..4) next to the automatic choice (0); prints one JSON object.  Used to fit the
cost model in choose_tile() (eigencuda_b200/csrc/ec_api.cu) and kept under profiles/."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigencuda_b200 as ec

dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
NAMES = {-1: "auto_nosk", 0: "auto", 1: "L128x128", 2: "S64x64", 3: "XS64x32"}


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


out = {"square": [], "batched": []}
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [256, 512, 768, 1024, 1536, 2048, 3072, 4096, 8192]
for n in sizes:
    A = torch.empty(n * n, dtype=torch.float64, device=dev)
    B = torch.empty(n * n, dtype=torch.float64, device=dev)
    Cm = torch.empty(n * n, dtype=torch.float64, device=dev)
    pipe.fill_uniform(1, 0, A.data_ptr(), n * n, 1)
    pipe.fill_uniform(1, 1, B.data_ptr(), n * n, 1)
    ref = None
    if n <= 2048:
        ref = (A.view(n, n).T.contiguous().T @ B.view(n, n).T.contiguous().T) if False else None
        # column-major n x n stored flat == (row-major view).T
        Am, Bm = A.view(n, n).T, B.view(n, n).T
        ref = (Am @ Bm)
    row = {"n": n}
    for tile in (-1, 0, 1, 2, 3):
        os.environ["EC_STREAM_K"] = "0" if tile < 0 else "1"
        if tile > 0:
            os.environ["EC_FORCE_TILE"] = str(tile)
        else:
            os.environ.pop("EC_FORCE_TILE", None)
        if tile == 1 and n < 128 or tile == 3 and n > 4096:
            continue
        fn = lambda: pipe.dgemm(0, 0, n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, Cm.data_ptr(), n)
        ms = timed(fn, 3 if n >= 8192 else 10)
        e = {"ms": round(ms, 4), "tflops": round(2.0 * n ** 3 / (ms * 1e-3) / 1e12, 2), "kernel": pipe.last_kernel()}
        if ref is not None:
            got = Cm.view(n, n).T
            e["rel_err"] = float(torch.linalg.norm(got - ref) / torch.linalg.norm(ref))
        row[NAMES[tile]] = e
    out["square"].append(row)
    del A, B, Cm
for n, batch in [(512, 64), (512, 1024), (512, 4096), (1000, 256), (256, 4096), (128, 8192)]:
    A = torch.empty(n * n * batch, dtype=torch.float64, device=dev)
    B = torch.empty(n * n, dtype=torch.float64, device=dev)
    Cm = torch.empty(n * n * batch, dtype=torch.float64, device=dev)
    pipe.fill_uniform(1, 1, A.data_ptr(), n * n, batch)
    pipe.fill_uniform(1, 0, B.data_ptr(), n * n, 1)
    row = {"n": n, "batch": batch}
    for tile in (-1, 0, 1, 2, 3):
        os.environ["EC_STREAM_K"] = "0" if tile < 0 else "1"
        if tile > 0:
            os.environ["EC_FORCE_TILE"] = str(tile)
        else:
            os.environ.pop("EC_FORCE_TILE", None)
        fn = lambda: pipe.dgemm_strided_batched(0, 0, n, n, n, 1.0, A.data_ptr(), n, n * n, B.data_ptr(), n, 0, 0.0,
                                                Cm.data_ptr(), n, n * n, batch)
        ms = timed(fn, 5)
        row[NAMES[tile]] = {"ms": round(ms, 4), "tflops": round(2.0 * n ** 3 * batch / (ms * 1e-3) / 1e12, 2),
                            "matrices_per_s": round(batch / (ms * 1e-3)), "kernel": pipe.last_kernel()}
    out["batched"].append(row)
    del A, B, Cm
os.environ.pop("EC_FORCE_TILE", None)
os.environ.pop("EC_STREAM_K", None)
print(json.dumps(out))
