"""Generates tests/golden/ref_cublas.npz by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref, i.e.
/root/reference/src/*.cc compiled against cuBLAS) on a B200:

    gpurun -- 'python tests/golden/make_ref_golden.py gpurun_out/ref_cublas.npz'
    cp gpurun_out/ref_cublas.npz tests/golden/

Inputs are the seeded synthetic matrices of SURVEY.md section 8(d) (regenerated, not stored);
outputs are what the reference's CudaPipeline::gemm (src/cudapipeline.cc:17-32) and its README
streaming loop (README.md:62-66) returned.  The CPU tests then pin the oracle against these."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as orc  # noqa: E402

CASES = [  # (name, m, k, n)
    ("square200", 200, 200, 200),   # the shape of src/tests/test_dot.cc:34-51
    ("rect", 96, 72, 40),
    ("odd", 33, 17, 29),
    ("square512", 512, 512, 512),
]
STREAM = ("stream", 6, 64, 64)      # 6 tensor elements 64x64 against a fixed 64x64


def main(out):
    ref = orc.Reference()
    data = {}
    for cid, (name, m, k, n) in enumerate(CASES):
        A = orc.fill_uniform(orc.SEED, 1000 + 2 * cid, (m, k))
        B = orc.fill_uniform(orc.SEED, 1001 + 2 * cid, (k, n))
        data[name] = ref.gemm(A, B)
    _, cnt, r, c = STREAM
    F = orc.fill_uniform(orc.SEED, 0, (c, c))
    ts = [orc.fill_uniform(orc.SEED, i + 1, (r, c)) for i in range(cnt)]
    res, _ = ref.tensor_stream(0, F, ts)
    data["stream_right"] = np.stack(res)
    res, _ = ref.tensor_stream(1, F, ts)
    data["stream_left"] = np.stack(res)
    data["roundtrip"] = ref.roundtrip(ts[0])
    np.savez_compressed(out, **data)
    print("wrote", out, {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ref_cublas.npz"))
