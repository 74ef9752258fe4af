"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle,
against the committed golden vectors and -- when oracle/_ref is present -- against the unmodified
reference (cuBLAS) on the same inputs.

Bar (BASELINE.json north_star): relative Frobenius error <= 1e-12 for the native FP64 path (the
reference's own tests use Eigen isApprox, whose default precision for double is 1e-12:
/root/reference/src/tests/test_dot.cc:31,50,87-89).  Byte-moving paths (H2D/D2H round trips, the
synthetic generator) must be bit-exact.  Full-size cases (BASELINE.json configs 2-4) are checked
through size-independent properties: Freivalds products, identity, linearity, sampled elements.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import eigencuda_b200 as ec
from oracle import oracle_py as orc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
TOL = 1e-12  # relative Frobenius, BASELINE.json north_star / Eigen isApprox default


@pytest.fixture(scope="module")
def pipe():
    if ec.count_available_gpus() < 1:
        pytest.fail("gpu tests need a B200; there is no CPU fallback")
    p = ec.CudaPipeline(0)
    yield p
    p.close()


@pytest.fixture(scope="module")
def ref():
    if not orc.Reference.available():
        pytest.skip("oracle/_ref not built (only built where /root/reference exists)")
    return orc.Reference()


def _mat(d):
    return np.asfortranarray(np.array(d["row_major"], dtype=np.float64).reshape(d["rows"], d["cols"]))


def rnd(seed, shape):
    return np.asfortranarray(np.random.default_rng(seed).uniform(-1, 1, shape))


# ---------------------------------------------------------------------------------------------
# the reference's own four test cases, through the mirrored API
# ---------------------------------------------------------------------------------------------
def test_create_cudamatrix_roundtrip(pipe):
    """src/tests/test_dot.cc:11-32: H2D in the ctor, D2H in the conversion, bit-exact."""
    g = json.load(open(os.path.join(GOLD, "readme_example.json")))
    A, B, X = _mat(g["A"]), _mat(g["tensor"][0]), _mat(g["expected"][0])
    cumatrix = ec.CudaMatrix(B, pipe.get_stream())
    tmp = cumatrix.to_eigen()
    assert np.array_equal(tmp, B)
    assert tmp.ravel(order="K").tolist() == g["column_major_bytes"]["B"]
    assert orc.is_approx(tmp @ A, X)
    assert (cumatrix.rows(), cumatrix.cols(), cumatrix.size()) == (3, 2, 6)


def test_matrix_multiplication_200(pipe, ref):
    """src/tests/test_dot.cc:34-51: 200x200 random, vs the CPU product with isApprox."""
    A, B = rnd(1, (200, 200)), rnd(2, (200, 200))
    cA, cB = ec.CudaMatrix(A, pipe.get_stream()), ec.CudaMatrix(B, pipe.get_stream())
    cC = ec.CudaMatrix(200, 200, pipe.get_stream())
    pipe.gemm(cA, cB, cC)
    got = cC.to_eigen()
    assert orc.is_approx(got, orc.cpu_product(A, B))            # "Eigen's CPU product"
    assert orc.rel_frobenius(got, orc.dgemm(A, B, flavour="ld")) <= TOL
    assert orc.rel_frobenius(got, ref.gemm(A, B)) <= TOL        # the reference's cuBLAS path


def test_right_matrix_multiplication_readme(pipe):
    """src/tests/test_dot.cc:53-90 / README.md:37-72: the streaming pattern, known answers."""
    g = json.load(open(os.path.join(GOLD, "readme_example.json")))
    A = _mat(g["A"])
    tensor = [_mat(t) for t in g["tensor"]]
    want = [_mat(x) for x in g["expected"]]
    cuma_A = ec.CudaMatrix(A, pipe.get_stream())
    cuma_B = ec.CudaMatrix(3, 2, pipe.get_stream())
    cuma_C = ec.CudaMatrix(3, 2, pipe.get_stream())
    results = []
    for i in range(3):
        cuma_B.copy_to_gpu(tensor[i])
        pipe.gemm(cuma_B, cuma_A, cuma_C)
        results.append(cuma_C.to_eigen())
    for r, x in zip(results, want):
        assert np.array_equal(r, x)          # small integers: exact
    # the same job as one overlapped call
    for r, x in zip(pipe.right_multiply(tensor, A), want):
        assert np.array_equal(r, x)


def test_wrong_shape_raises_before_device_work(pipe):
    """src/tests/test_dot.cc:92-103."""
    cA = ec.CudaMatrix(rnd(3, (2, 2)), pipe.get_stream())
    cB = ec.CudaMatrix(rnd(4, (5, 5)), pipe.get_stream())
    cC = ec.CudaMatrix(2, 5, pipe.get_stream())
    before = pipe.launch_count()
    with pytest.raises(RuntimeError) as e:
        pipe.gemm(cA, cB, cC)
    assert e.value.code == 1 and "Shape mismatch" in str(e.value)
    assert pipe.launch_count() == before
    with pytest.raises(RuntimeError):
        pipe.right_multiply([rnd(5, (3, 2))], rnd(6, (5, 5)))


def test_cpp_compat_binary(tmp_path):
    """Caller code written like the reference's tests, compiled against include/*.hpp, runs green."""
    exe = tmp_path / "test_dot_compat"
    cmd = ["g++", "-std=c++14", "-O2", f"-I{ROOT}/include", f"-I{ROOT}/oracle/eigen_standin",
           "-I/usr/local/cuda/include", f"{ROOT}/tests/cpp/test_dot_compat.cc", "-o", str(exe),
           f"-L{ROOT}/eigencuda_b200/lib", "-leigencuda_b200", f"-Wl,-rpath,{ROOT}/eigencuda_b200/lib"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ALL PASSED" in res.stdout, res.stdout + res.stderr


# ---------------------------------------------------------------------------------------------
# golden fixture captured from the unmodified reference
# ---------------------------------------------------------------------------------------------
def test_against_committed_reference_outputs(pipe):
    path = os.path.join(GOLD, "ref_cublas.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ref_cublas.npz not generated yet")
    from tests.golden.make_ref_golden import CASES, STREAM
    z = np.load(path)
    for cid, (name, m, k, n) in enumerate(CASES):
        A = orc.fill_uniform(orc.SEED, 1000 + 2 * cid, (m, k))
        B = orc.fill_uniform(orc.SEED, 1001 + 2 * cid, (k, n))
        cA, cB = ec.CudaMatrix(A, pipe.get_stream()), ec.CudaMatrix(B, pipe.get_stream())
        cC = ec.CudaMatrix(m, n, pipe.get_stream())
        pipe.gemm(cA, cB, cC)
        assert orc.rel_frobenius(cC.to_eigen(), z[name]) <= TOL, name
    _, cnt, r, c = STREAM
    F = orc.fill_uniform(orc.SEED, 0, (c, c))
    ts = [orc.fill_uniform(orc.SEED, i + 1, (r, c)) for i in range(cnt)]
    for got, want in zip(pipe.right_multiply(ts, F), z["stream_right"]):
        assert orc.rel_frobenius(got, want) <= TOL
    for got, want in zip(pipe.left_multiply(F, ts), z["stream_left"]):
        assert orc.rel_frobenius(got, want) <= TOL


# ---------------------------------------------------------------------------------------------
# the DGEMM kernels: every op(A)/op(B), ragged shapes, leading dimensions, alpha/beta
# ---------------------------------------------------------------------------------------------
SHAPES = [  # (m, n, k)
    (128, 128, 16), (128, 128, 128), (256, 384, 512), (130, 258, 70), (200, 200, 200),
    (64, 64, 64), (520, 136, 1030), (32, 1024, 48), (1024, 40, 24), (40, 72, 96),
    # skinny: narrower than the smallest tile in one dimension, TMA boxes hang over the edge
    (4096, 16, 128), (16, 4096, 128), (2048, 8, 64), (8, 8, 4096), (1000, 24, 12),
]


def run_dgemm(pipe, ta, tb, m, n, k, alpha, beta, lda_pad=0, ldb_pad=0, ldc_pad=0, seed=0):
    a_shape = (k, m) if ta else (m, k)
    b_shape = (n, k) if tb else (k, n)
    Afull = rnd(seed + 1, (a_shape[0] + lda_pad, a_shape[1]))
    Bfull = rnd(seed + 2, (b_shape[0] + ldb_pad, b_shape[1]))
    Cfull = rnd(seed + 3, (m + ldc_pad, n))
    dA, dB = ec.CudaMatrix(Afull, pipe.get_stream()), ec.CudaMatrix(Bfull, pipe.get_stream())
    dC = ec.CudaMatrix(Cfull, pipe.get_stream())
    pipe.dgemm(int(ta), int(tb), m, n, k, alpha, dA.data(), Afull.shape[0], dB.data(),
               Bfull.shape[0], beta, dC.data(), Cfull.shape[0])
    got_full = dC.to_eigen()
    A, B = Afull[:a_shape[0]], Bfull[:b_shape[0]]
    want = orc.dgemm(A, B, ta, tb, alpha, beta, Cin=Cfull[:m], flavour="ld")
    # rows of C beyond m (the ldc padding) must be untouched
    assert np.array_equal(got_full[m:], Cfull[m:])
    return got_full[:m], want


@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
@pytest.mark.parametrize("shape", SHAPES)
def test_dgemm_tma_kernel_all_layouts(pipe, ta, tb, shape):
    m, n, k = shape
    got, want = run_dgemm(pipe, ta, tb, m, n, k, 1.0, 0.0, seed=m + n + k)
    assert pipe.last_kernel().startswith("tma_dmma_" + ("T" if ta else "N") + ("T" if tb else "N"))
    assert orc.rel_frobenius(got, want) <= TOL


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_dgemm_alpha_beta_and_padding(pipe, ta, tb):
    got, want = run_dgemm(pipe, ta, tb, 200, 136, 88, -0.5, 2.25, lda_pad=6, ldb_pad=2, ldc_pad=4, seed=9)
    assert pipe.last_kernel().startswith("tma_dmma")
    assert orc.rel_frobenius(got, want) <= TOL


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("shape", [(3, 2, 2), (1, 1, 1), (7, 5, 3), (65, 33, 17), (129, 131, 67), (31, 200, 257)])
def test_dgemm_simt_kernel_odd_shapes(pipe, ta, tb, shape):
    """Odd leading dimensions cannot be TMA-addressed (16-byte pitch rule): generic kernel."""
    m, n, k = shape
    got, want = run_dgemm(pipe, ta, tb, m, n, k, 1.5, -0.75, lda_pad=0, ldb_pad=0, ldc_pad=1, seed=5)
    if (k if ta else m) % 2 or (n if tb else k) % 2 or m * n * k < 64 ** 3:
        assert pipe.last_kernel() == "simt"
    assert orc.rel_frobenius(got, want) <= TOL


@pytest.mark.parametrize("tile", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_every_tile_configuration(pipe, tile, ta, tb, monkeypatch):
    """The five CTA tile shapes (128x128, 64x64, 64x32, 32x32, 128x16) compute the same product."""
    monkeypatch.setenv("EC_FORCE_TILE", str(tile))
    for (m, n, k) in [(256, 192, 144), (130, 70, 50), (520, 264, 1030)]:
        got, want = run_dgemm(pipe, ta, tb, m, n, k, 1.25, -0.5, seed=tile + m)
        suffix = {1: "", 2: "_s", 3: "_xs", 4: "_t", 5: "_n"}[tile]
        name = "tma_dmma_" + ("T" if ta else "N") + ("T" if tb else "N") + suffix
        assert pipe.last_kernel() in (name, name + "_sk")
        assert orc.rel_frobenius(got, want) <= TOL


@pytest.mark.parametrize("tile", [1, 2, 3, 4])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, True)])
def test_stream_k_schedule(pipe, tile, ta, tb, monkeypatch):
    """Shapes whose tile count does not fill the last wave take the stream-K schedule: partial
    sums cross CTAs through the workspace, the owner adds them in CTA order.  Same bar, and the
    result is bit-reproducible from run to run (the fix-up order is fixed)."""
    monkeypatch.setenv("EC_FORCE_TILE", str(tile))
    monkeypatch.setenv("EC_STREAM_K", "2")     # stream-K whenever feasible, not only when cheaper
    for (m, n, k) in [(1280, 1280, 512), (900, 1100, 2000), (1536, 1024, 160)]:
        got, want = run_dgemm(pipe, ta, tb, m, n, k, 1.0, 0.5, seed=tile)
        assert pipe.last_kernel().endswith("_sk"), (pipe.last_kernel(), m, n, k)
        assert orc.rel_frobenius(got, want) <= TOL
        again, _ = run_dgemm(pipe, ta, tb, m, n, k, 1.0, 0.5, seed=tile)
        assert np.array_equal(got, again)
    monkeypatch.setenv("EC_STREAM_K", "0")
    got0, want = run_dgemm(pipe, ta, tb, 1280, 1280, 512, 1.0, 0.5, seed=tile)
    assert not pipe.last_kernel().endswith("_sk")
    assert orc.rel_frobenius(got0, want) <= TOL


@pytest.mark.parametrize("tile", [1, 2, 3, 4])
@pytest.mark.parametrize("shape", [(64, 64, 200000), (128, 96, 120000), (200, 40, 70000)])
def test_stream_k_deep_split_two_level_sum(pipe, tile, shape, monkeypatch):
    """Inner-product shapes: a handful of output tiles, K in the hundred thousands, so every CTA
    of the grid contributes to the same tile.  Beyond 8 contributors the fix-up sums in two levels
    (groups of 16 ranks); same bar against the long-double oracle, bit-reproducible run to run."""
    monkeypatch.setenv("EC_FORCE_TILE", str(tile))
    m, n, k = shape
    A, B = rnd(k, (m, k)), rnd(k + 1, (k, n))
    dA, dB = ec.CudaMatrix(A, pipe.get_stream()), ec.CudaMatrix(B, pipe.get_stream())
    dC = ec.CudaMatrix(m, n, pipe.get_stream())
    pipe.gemm(dA, dB, dC)
    assert pipe.last_kernel().endswith("_sk"), pipe.last_kernel()
    got = dC.to_eigen()
    want = orc.dgemm(A, B, flavour="ld")
    assert orc.rel_frobenius(got, want) <= TOL
    pipe.gemm(dA, dB, dC)
    assert np.array_equal(dC.to_eigen(), got)
    # transposed operands (K-major A, MN-major B) through the raw entry
    dAt = ec.CudaMatrix(np.asfortranarray(A.T), pipe.get_stream())
    dBt = ec.CudaMatrix(np.asfortranarray(B.T), pipe.get_stream())
    pipe.dgemm(1, 1, m, n, k, 1.0, dAt.data(), k, dBt.data(), n, 0.0, dC.data(), m)
    assert pipe.last_kernel().endswith("_sk"), pipe.last_kernel()
    assert orc.rel_frobenius(dC.to_eigen(), want) <= TOL


def test_slot_release_race_regression(pipe, monkeypatch):
    """Regression: with several CTAs per SM, a consumer's empty-barrier arrive used to be able
    to overtake its own in-flight LDS (queued behind another CTA's epilogue stores), letting the
    TMA refill land under the loads -- rare corrupted 16-row fragments.  Every tile configuration
    must reproduce the 128x128 configuration's result BIT FOR BIT on a long batched run (the
    per-element summation order does not depend on the tile shape)."""
    n, batch = 512, 600
    dT = ec.CudaMatrix(n, n * batch, pipe.get_stream())
    dF = ec.CudaMatrix(n, n, pipe.get_stream())
    pipe.fill_uniform(orc.SEED, 1, dT.data(), n * n, batch)
    pipe.fill_uniform(orc.SEED, 0, dF.data(), n * n, 1)
    monkeypatch.setenv("EC_STREAM_K", "0")

    def run(tile):
        monkeypatch.setenv("EC_FORCE_TILE", str(tile))
        dR = ec.CudaMatrix(n, n * batch, pipe.get_stream())
        pipe.dgemm_strided_batched(0, 0, n, n, n, 1.0, dT.data(), n, n * n, dF.data(), n, 0, 0.0, dR.data(), n, n * n, batch)
        return dR.to_eigen()

    want = run(1)
    F = orc.fill_uniform(orc.SEED, 0, (n, n))
    assert orc.rel_frobenius(want[:, :n], orc.cpu_product(orc.fill_uniform(orc.SEED, 1, (n, n)), F)) <= TOL
    for tile in (2, 3, 4, 5):
        for _ in range(4):
            assert np.array_equal(run(tile), want), f"tile configuration {tile} diverged"


def test_stream_k_batched(pipe, monkeypatch):
    """The tensor-stream shape with a batch that leaves a partial last wave."""
    monkeypatch.setenv("EC_STREAM_K", "2")
    n, batch = 256, 41
    dT = ec.CudaMatrix(n, n * batch, pipe.get_stream())
    dF = ec.CudaMatrix(n, n, pipe.get_stream())
    dR = ec.CudaMatrix(n, n * batch, pipe.get_stream())
    pipe.fill_uniform(orc.SEED, 1, dT.data(), n * n, batch)
    pipe.fill_uniform(orc.SEED, 0, dF.data(), n * n, 1)
    pipe.dgemm_strided_batched(0, 0, n, n, n, 1.0, dT.data(), n, n * n, dF.data(), n, 0, 0.0, dR.data(), n, n * n, batch)
    assert pipe.last_kernel().endswith("_sk"), pipe.last_kernel()
    got = dR.to_eigen().reshape(n, n, batch, order="F")
    F = orc.fill_uniform(orc.SEED, 0, (n, n))
    for i in range(batch):
        assert orc.rel_frobenius(got[:, :, i], orc.cpu_product(orc.fill_uniform(orc.SEED, i + 1, (n, n)), F)) <= TOL


def test_dynamic_tile_schedule_matches_static(monkeypatch):
    """Whole tiles claimed from a counter (the default for tiles of 8+ k-iterations) give the same
    bits as the static round-robin schedule: which CTA computes a tile does not change its sums."""
    shapes = [(2048, 2048, 512, 1), (1000, 1000, 1000, 7), (512, 512, 512, 64), (640, 384, 2048, 3)]
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("EC_DYNAMIC_TILES", mode)
        monkeypatch.setenv("EC_STREAM_K", "0")
        p = ec.CudaPipeline(0)
        for (m, n, k, batch) in shapes:
            A, B = rnd(31, (m * batch, k)), rnd(32, (k, n))
            dA, dB = ec.CudaMatrix(A, p.get_stream()), ec.CudaMatrix(B, p.get_stream())
            dC = ec.CudaMatrix(m * batch, n, p.get_stream())
            for _ in range(3):          # repeated launches: the claim counter is never reset
                p.dgemm(0, 0, m * batch, n, k, 1.0, dA.data(), m * batch, dB.data(), k, 0.0, dC.data(), m * batch)
            got = dC.to_eigen()
            outs.setdefault((m, n, k, batch), []).append(got)
            assert orc.rel_frobenius(got, orc.dgemm(A, B, flavour="fast")) <= TOL
        p.close()
    for key, (a, b) in outs.items():
        assert np.array_equal(a, b), key


def test_fused_tile_turnaround_matches_plain_epilogue(monkeypatch):
    """The 128x128 tile stores a finished tile underneath the first k-iteration of the CTA's next
    tile (row-block pair by row-block pair: store, clear, DMMA).  Every accumulator still sees its
    k-steps in the same order, so the bits must equal the plain epilogue's (EC_FUSE_EPILOGUE=0) --
    across all four operand layouts, multi-tile CTAs (persistent grid, stream-K region, dynamic
    claims), edge tiles next to interior ones (deferred as well, by the kernel instantiation with predicated stores), K of one k-iteration (the fused iteration is also the
    last) and beta != 0 (never deferred)."""
    shapes = [  # (transa, transb, m, n, k, batch, beta, alpha); alpha == 1 skips the multiply in the fused stores
        (0, 0, 512, 512, 512, 40, 0.0, 1.0), (1, 0, 512, 512, 512, 40, 0.0, 1.25), (0, 1, 512, 512, 512, 40, 0.0, 1.25),
        (1, 1, 512, 512, 512, 40, 0.0, 1.0), (0, 0, 4096, 4096, 256, 1, 0.0, 1.25), (0, 0, 4000, 3000, 100, 1, 0.0, 1.0),
        (0, 0, 2048, 2048, 16, 2, 0.0, 1.0), (0, 0, 2048, 2048, 8, 2, 0.0, 1.25), (0, 0, 2304, 2304, 528, 1, 0.5, 1.25),
        (1, 0, 1536, 2560, 48, 3, 0.0, 1.0), (0, 0, 3072, 3072, 384, 1, 0.0, 1.0),
        # partial tiles: the edge instantiation defers them too (predicated stores), odd M / odd N included
        (0, 0, 1000, 1000, 1000, 6, 0.0, 1.0), (1, 0, 1001, 999, 320, 1, 0.0, 1.25), (0, 1, 1000, 999, 200, 2, 0.0, 1.0),
        (1, 1, 130, 4100, 64, 1, 0.0, 1.0)]
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("EC_FUSE_EPILOGUE", mode)
        monkeypatch.setenv("EC_FORCE_TILE", "1")
        p = ec.CudaPipeline(0)
        for key in shapes:
            ta, tb, m, n, k, batch, beta, alpha = key
            A = rnd(51, ((k if ta else m) * batch, (m if ta else k)))
            B = rnd(52, ((n if tb else k), (k if tb else n)))
            C0 = rnd(53, (m * batch, n))
            lda, ldb = A.shape[0], B.shape[0]
            dA, dB = ec.CudaMatrix(A, p.get_stream()), ec.CudaMatrix(B, p.get_stream())
            dC = ec.CudaMatrix(C0, p.get_stream())
            if batch == 1:
                p.dgemm(ta, tb, m, n, k, alpha, dA.data(), lda, dB.data(), ldb, beta, dC.data(), m)
                got = dC.to_eigen()
                opA = A.T if ta else A
                opB = B.T if tb else B
                ref = alpha * orc.dgemm(np.asfortranarray(opA), np.asfortranarray(opB), flavour="fast") + beta * C0
                assert orc.rel_frobenius(got, ref) <= TOL, key
            else:
                # problem i occupies rows [i*rows_a, (i+1)*rows_a) of the stacked A and rows [i*m, (i+1)*m) of C
                rows_a = k if ta else m
                p.dgemm_strided_batched(ta, tb, m, n, k, alpha, dA.data(), lda, rows_a, dB.data(), ldb, 0, beta,
                                        dC.data(), m * batch, m, batch)
                got = dC.to_eigen()
                opB = B.T if tb else B
                for i in (0, batch - 1):
                    Ai = A[i * rows_a:(i + 1) * rows_a, :]
                    opA = Ai.T if ta else Ai
                    ref = alpha * orc.dgemm(np.asfortranarray(opA), np.asfortranarray(opB), flavour="fast") \
                        + beta * C0[i * m:(i + 1) * m, :]
                    assert orc.rel_frobenius(got[i * m:(i + 1) * m, :], ref) <= TOL, (key, i)
            outs.setdefault(key, []).append(got)
        p.close()
    for key, (a, b) in outs.items():
        assert np.array_equal(a, b), key


def test_shared_gpu_is_detected_and_schedules_adapt(monkeypatch):
    """a statically partitioned grid (stream-K above all) runs at the pace of the CTA
    that gets its SM last.  With another kernel holding 24 of the 148 SMs, a static grid takes two
    rounds (the CTAs that found no SM start when the first ones leave: 2.0x); with tiles claimed
    dynamically the first product already does better, it reports the start skew of its CTAs, and
    from the next launch on the dispatcher drops the stream-K region too (1.3x, the ideal being
    148/124 = 1.19x).  Launches are plain (no cooperative launch), results unchanged."""
    import time
    import torch
    n = 4096
    side = torch.cuda.Stream()
    A, B = rnd(41, (n, n)), rnd(42, (n, n))

    def measure(dynamic):
        monkeypatch.setenv("EC_DYNAMIC_TILES", "1" if dynamic else "0")
        p = ec.CudaPipeline(0)
        dA, dB = ec.CudaMatrix(A, p.get_stream()), ec.CudaMatrix(B, p.get_stream())
        dC = ec.CudaMatrix(n, n, p.get_stream())
        run = lambda: p.dgemm(0, 0, n, n, n, 1.0, dA.data(), n, dB.data(), n, 0.0, dC.data(), n)
        for _ in range(2):
            run()
        p.synchronize()
        assert p.last_kernel().endswith("_sk") and p.contention_events() == 0
        quiet = dC.to_eigen()
        times, kernels = [], []
        for _ in range(3):
            ec.debug_hold_sms(0, side.cuda_stream, 24, 40000)     # 24 SMs taken for 40 ms
            time.sleep(0.003)                                     # let it become resident
            t0 = time.perf_counter()
            run()
            p.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
            kernels.append(p.last_kernel())
            side.synchronize()
            assert np.array_equal(dC.to_eigen(), quiet) or orc.rel_frobenius(dC.to_eigen(), quiet) <= 1e-14
        events = p.contention_events()
        p.close()
        return times, kernels, events, quiet

    t_static, k_static, e_static, ref = measure(False)
    t_dyn, k_dyn, e_dyn, got = measure(True)
    assert e_static == 0 and all(k.endswith("_sk") for k in k_static)
    assert e_dyn >= 1 and k_dyn[0].endswith("_sk") and not k_dyn[-1].endswith("_sk"), (k_dyn, e_dyn)
    assert t_dyn[-1] < 0.85 * t_static[-1], (t_static, t_dyn)       # measured 4.95 ms against 7.57 ms
    assert orc.rel_frobenius(got[:64], orc.dgemm(A[:64], B, flavour="fast")) <= TOL
    assert orc.rel_frobenius(got, ref) <= 1e-14


def test_one_cta_per_tile_schedule(pipe):
    """The non-persistent launch (one CTA per tile, for long-K tiles on a shared GPU) computes the
    same bits as the persistent grid."""
    A, B = rnd(1, (1300, 1100)), rnd(2, (1100, 1800))
    dA, dB = ec.CudaMatrix(A, pipe.get_stream()), ec.CudaMatrix(B, pipe.get_stream())
    dC = ec.CudaMatrix(1300, 1800, pipe.get_stream())
    pipe.set_schedule(stream_k=False, one_cta_per_tile=False)
    pipe.gemm(dA, dB, dC)
    want = dC.to_eigen()
    pipe.set_schedule(stream_k=False, one_cta_per_tile=True)
    pipe.gemm(dA, dB, dC)
    got = dC.to_eigen()
    pipe.set_schedule(stream_k=True, one_cta_per_tile=False)
    assert np.array_equal(got, want)
    assert orc.rel_frobenius(got, orc.cpu_product(A, B)) <= TOL


@pytest.mark.parametrize("kind", ["right", "left", "left_t", "basis"])
@pytest.mark.parametrize("n", [8, 16, 24, 32, 48])
def test_tensor_of_tiny_matrices_resident(pipe, kind, n):
    """A tensor of matrices smaller than the smallest tile: the batch as a whole decides the
    kernel (one HBM-bound job), each matrix is one partly empty tile; every stream kind."""
    count = 777
    F = orc.fill_uniform(orc.SEED, 0, (n, n))
    ts = [orc.fill_uniform(orc.SEED, i + 1, (n, n)) for i in range(count)]
    code = {"right": ec.EC_STREAM_RIGHT, "left": ec.EC_STREAM_LEFT, "left_t": ec.EC_STREAM_LEFT_T,
            "basis": ec.EC_STREAM_BASIS}[kind]
    res = pipe.tensor_stream(code, F, ts)
    assert pipe.last_kernel().startswith("tma_dmma"), pipe.last_kernel()
    want = orc.tensor_stream(kind, F, ts, flavour="fast")
    for r, w in zip(res, want):
        assert orc.rel_frobenius(r, w) <= TOL


def test_beta_zero_ignores_nan_in_c(pipe):
    A, B = rnd(1, (160, 96)), rnd(2, (96, 144))
    dA, dB = ec.CudaMatrix(A, pipe.get_stream()), ec.CudaMatrix(B, pipe.get_stream())
    dC = ec.CudaMatrix(np.full((160, 144), np.nan, order="F"), pipe.get_stream())
    pipe.gemm(dA, dB, dC)
    got = dC.to_eigen()
    assert np.isfinite(got).all()
    assert orc.rel_frobenius(got, orc.dgemm(A, B, flavour="ld")) <= TOL


def test_k_zero_and_alpha_zero_scale_c(pipe):
    Cin = rnd(7, (40, 24))
    dC = ec.CudaMatrix(Cin, pipe.get_stream())
    dA = ec.CudaMatrix(rnd(8, (40, 8)), pipe.get_stream())
    dB = ec.CudaMatrix(rnd(9, (8, 24)), pipe.get_stream())
    pipe.dgemm(0, 0, 40, 24, 8, 0.0, dA.data(), 40, dB.data(), 8, 3.0, dC.data(), 40)
    assert np.array_equal(dC.to_eigen(), 3.0 * Cin)
    pipe.dgemm(0, 0, 40, 24, 0, 1.0, 0, 40, 0, 1, 0.0, dC.data(), 40)
    assert np.array_equal(dC.to_eigen(), np.zeros((40, 24)))


def test_empty_products_are_noops(pipe):
    before = pipe.launch_count()
    pipe.dgemm(0, 0, 0, 5, 3, 1.0, 0, 1, 0, 3, 0.0, 0, 1)
    pipe.dgemm(0, 0, 5, 0, 3, 1.0, 0, 5, 0, 3, 0.0, 0, 5)
    assert pipe.launch_count() == before
    assert pipe.right_multiply([], np.eye(2)) == []


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True)])
def test_strided_batched_matches_loop(pipe, ta, tb):
    m, n, k, batch = 136, 72, 200, 5
    a_shape = (k, m) if ta else (m, k)
    b_shape = (n, k) if tb else (k, n)
    As = np.asfortranarray(np.stack([rnd(i, a_shape) for i in range(batch)], axis=2))
    Bf = rnd(99, b_shape)                      # shared operand: stride 0
    dA = ec.CudaMatrix(As.reshape(a_shape[0], -1, order="F"), pipe.get_stream())
    dB = ec.CudaMatrix(Bf, pipe.get_stream())
    dC = ec.CudaMatrix(m, n * batch, pipe.get_stream())
    pipe.dgemm_strided_batched(int(ta), int(tb), m, n, k, 1.0, dA.data(), a_shape[0],
                               a_shape[0] * a_shape[1], dB.data(), b_shape[0], 0, 0.0, dC.data(), m,
                               m * n, batch)
    assert pipe.last_kernel().startswith("tma_dmma")
    got = dC.to_eigen().reshape(m, n, batch, order="F")
    for i in range(batch):
        want = orc.dgemm(As[:, :, i], Bf, ta, tb, flavour="ld")
        assert orc.rel_frobenius(got[:, :, i], want) <= TOL


# ---------------------------------------------------------------------------------------------
# memory subsystem: round trips are bit-exact, including the staged large-transfer paths
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 1), (3, 2), (256, 255), (512, 512), (1031, 257), (517, 1100),
                                   (1000, 1000), (1537, 911)])
def test_upload_download_bit_exact(pipe, shape):
    A = rnd(shape[0], shape)
    dA = ec.CudaMatrix(A, pipe.get_stream())
    assert np.array_equal(dA.to_eigen(), A)
    B = rnd(shape[1] + 1, shape)
    dA.copy_to_gpu(B)
    B_copy = B.copy(order="F")
    B[:] = 0.0                       # source is reusable as soon as copy_to_gpu returns
    assert np.array_equal(dA.to_eigen(), B_copy)


def test_matrices_constructed_inside_the_loop(pipe):
    """Callers that build their CudaMatrix objects per iteration (allocation from the stream-ordered
    pool, release in stream order behind the product that still reads them): results identical."""
    n = 96
    F = rnd(1, (n, n))
    dF = ec.CudaMatrix(F, pipe.get_stream())
    for i in range(300):
        T = rnd(100 + i, (n, n))
        dT = ec.CudaMatrix(T, pipe.get_stream())
        dR = ec.CudaMatrix(n, n, pipe.get_stream())
        pipe.gemm(dT, dF, dR)
        dT.close()                       # freed while the product may still be running
        if i % 37 == 0:
            assert orc.rel_frobenius(dR.to_eigen(), orc.dgemm(T, F)) <= TOL
        dR.close()


def test_matrix_may_outlive_its_pipeline():
    """The reference's deleter is a plain cudaFree (include/cudamatrix.hpp:46), so destroying the
    pipeline first is tolerated there; here too, although the buffer belongs to the stream's pool."""
    p = ec.CudaPipeline()
    A = rnd(3, (64, 48))
    dA = ec.CudaMatrix(A, p.get_stream())
    assert np.array_equal(dA.to_eigen(), A)
    p.close()
    dA.close()
    p2 = ec.CudaPipeline()
    dB = ec.CudaMatrix(A, p2.get_stream())
    assert np.array_equal(dB.to_eigen(), A)


def test_copy_to_gpu_rejects_oversized_source(pipe):
    d = ec.CudaMatrix(3, 2, pipe.get_stream())
    with pytest.raises(RuntimeError):
        d.copy_to_gpu(np.zeros((4, 2)))


def test_out_of_memory_raises_runtime_error(pipe):
    """src/cudamatrix.cc:63-79: a request beyond free device memory throws std::runtime_error."""
    with pytest.raises(RuntimeError) as e:
        ec.CudaMatrix(1 << 20, 1 << 20, pipe.get_stream())     # 8 TiB
    assert e.value.code == 2 and "not enough memory" in str(e.value)


def test_fill_uniform_matches_oracle_bit_for_bit(pipe):
    n, batch = 1000, 3
    d = ec.CudaMatrix(n, batch, pipe.get_stream())
    pipe.fill_uniform(orc.SEED, 5, d.data(), n, batch)
    got = d.to_eigen()
    for b in range(batch):
        assert np.array_equal(got[:, b], orc.fill_uniform(orc.SEED, 5 + b, (n,)))
    gold = json.load(open(os.path.join(GOLD, "uniform_kat.json")))
    e = gold["values"][0]
    d1 = ec.CudaMatrix(1, 1, pipe.get_stream())
    pipe.fill_uniform(gold["seed"], e["matrix_id"], d1.data(), 1, 1)
    assert d1.to_eigen()[0, 0] == float.fromhex(e["hex"])


def test_fill_uniform_block_matches_global_generator(pipe):
    """Sub-blocks generated by global index (2D-partitioned DGEMM operands) are bit-identical to
    the corresponding window of the whole matrix from the CPU generator."""
    gl_rows, gl_cols = 300, 200
    whole = orc.fill_uniform(orc.SEED, 9, (gl_rows, gl_cols))
    r0, c0, rows, cols, ld = 37, 50, 120, 64, 128
    d = ec.CudaMatrix(np.zeros((ld, cols), order="F"), pipe.get_stream())
    pipe.fill_uniform_block(orc.SEED, 9, d.data(), rows, cols, ld, r0, c0, gl_rows)
    got = d.to_eigen()
    assert np.array_equal(got[:rows], whole[r0:r0 + rows, c0:c0 + cols])
    assert np.array_equal(got[rows:], np.zeros((ld - rows, cols)))
    row = orc.uniform_strided(orc.SEED, 9, r0, gl_rows, gl_cols)
    assert np.array_equal(row, whole[r0, :])


# ---------------------------------------------------------------------------------------------
# tensor stream: every kind, pageable (staged) and pinned (direct DMA) host memory
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["right", "left", "left_t", "basis"])
@pytest.mark.parametrize("pinned", [False, True])
def test_tensor_stream_kinds(pipe, kind, pinned):
    n, count = 160, 23
    F = orc.fill_uniform(orc.SEED, 0, (n, n))
    bufs = []
    if pinned:
        pin = ec.PinnedBuffer((n, n, count))
        pout = ec.PinnedBuffer((n, n, count))
        bufs = [pin, pout]
        ts = [pin.array[:, :, i] for i in range(count)]
        for i, t in enumerate(ts):
            t[:] = orc.fill_uniform(orc.SEED, i + 1, (n, n))
        outs = [pout.array[:, :, i] for i in range(count)]
    else:
        ts = [orc.fill_uniform(orc.SEED, i + 1, (n, n)) for i in range(count)]
        outs = None
    pipe.set_stream_config(chunk_matrices=4, ring_depth=3, staging_threads=4)   # 6 chunks, ragged tail
    code = {"right": ec.EC_STREAM_RIGHT, "left": ec.EC_STREAM_LEFT, "left_t": ec.EC_STREAM_LEFT_T,
            "basis": ec.EC_STREAM_BASIS}[kind]
    res = pipe.tensor_stream(code, F, ts, results=outs)
    want = orc.tensor_stream(kind, F, ts, flavour="fast")
    for r, w in zip(res, want):
        assert orc.rel_frobenius(r, w) <= TOL
    pipe.set_stream_config(0, 0, 0)
    for b in bufs:
        b.close()


# ---------------------------------------------------------------------------------------------
# the multi-GPU tensor scheduler (ec_tensor_stream_run_multi): one vector, a device list
# ---------------------------------------------------------------------------------------------
def _device_sets():
    g = ec.count_available_gpus()
    sets = [[0]]
    k = 2
    while k <= g:
        sets.append(list(range(k)))
        k *= 2
    if g >= 2:
        sets.append([1, 0])
    sets.append("all")
    sets.append("auto")
    return sets


@pytest.mark.parametrize("kind", ["right", "left", "left_t", "basis"])
@pytest.mark.parametrize("pinned", [False, True])
def test_multi_gpu_tensor_stream_matches_oracle_and_single_gpu(pipe, kind, pinned):
    """Sharding over 1, 2, 4, ... devices (as many as the box has): results within 1e-12 of the
    oracle and BIT-identical to the single-device call; ragged shards; pinned and pageable."""
    n, count = 144, 29
    F = orc.fill_uniform(orc.SEED, 0, (n, n))
    bufs = []
    if pinned:
        pin, pout = ec.PinnedBuffer((n, n, count)), ec.PinnedBuffer((n, n, count))
        bufs = [pin, pout]
        ts = [pin.array[:, :, i] for i in range(count)]
        for i, t in enumerate(ts):
            t[:] = orc.fill_uniform(orc.SEED, i + 1, (n, n))
        outs = [pout.array[:, :, i] for i in range(count)]
    else:
        ts = [orc.fill_uniform(orc.SEED, i + 1, (n, n)) for i in range(count)]
        outs = None
    code = {"right": ec.EC_STREAM_RIGHT, "left": ec.EC_STREAM_LEFT, "left_t": ec.EC_STREAM_LEFT_T,
            "basis": ec.EC_STREAM_BASIS}[kind]
    pipe.set_stream_config(chunk_matrices=3, ring_depth=3, staging_threads=4)
    single = [r.copy() for r in pipe.tensor_stream(code, F, ts)]
    want = orc.tensor_stream(kind, F, ts, flavour="fast")
    for devices in _device_sets():
        if outs is not None:
            for o in outs:
                o[:] = np.nan
        res = pipe.tensor_stream(code, F, ts, results=outs, devices=devices)
        assert len(res) == count
        for r, w, s1 in zip(res, want, single):
            assert orc.rel_frobenius(r, w) <= TOL, devices
            assert np.array_equal(r, s1), devices
    pipe.set_stream_config(0, 0, 0)
    for b in bufs:
        b.close()


def test_multi_gpu_readme_example_and_edge_cases(pipe):
    g = json.load(open(os.path.join(GOLD, "readme_example.json")))
    A = _mat(g["A"])
    ts = [_mat(t) for t in g["tensor"]]
    for devices in _device_sets():
        res = pipe.right_multiply(ts, A, devices=devices)
        for r, x in zip(res, g["expected"]):
            assert orc.is_approx(r, _mat(x)), devices
        # fewer matrices than devices, and none at all
        one = pipe.right_multiply(ts[:1], A, devices=devices)
        assert len(one) == 1 and orc.is_approx(one[0], _mat(g["expected"][0]))
        assert pipe.right_multiply([], A, devices=devices) == []
    before = pipe.launch_count()
    with pytest.raises(RuntimeError) as e:
        pipe.right_multiply([rnd(5, (3, 2))], rnd(6, (5, 5)), devices="all")
    assert e.value.code == 1 and pipe.launch_count() == before     # before any device work
    with pytest.raises(RuntimeError):
        pipe.right_multiply(ts, A, devices=[0, 0])
    with pytest.raises(RuntimeError):
        pipe.right_multiply(ts, A, devices=[ec.count_available_gpus()])


def test_multi_gpu_counts_launches_and_keeps_its_state(pipe):
    n, count = 256, 16
    F = orc.fill_uniform(orc.SEED, 0, (n, n))
    ts = [orc.fill_uniform(orc.SEED, i + 1, (n, n)) for i in range(count)]
    pipe.set_stream_config(chunk_matrices=4, ring_depth=2, staging_threads=2)
    pipe.right_multiply(ts, F, devices="all")
    l0 = pipe.launch_count()
    res = pipe.right_multiply(ts, F, devices="all")
    g = ec.count_available_gpus()
    chunks = sum(-(-(ec.plan_shard(count, g, i)[1] - ec.plan_shard(count, g, i)[0]) // 4) for i in range(g))
    assert pipe.launch_count() - l0 == chunks          # one batched-DGEMM launch per chunk per device
    assert pipe.last_kernel().startswith("tma_dmma")
    for t, r in zip(ts, res):
        assert orc.rel_frobenius(r, orc.dgemm(t, F, flavour="ld")) <= TOL
    pipe.set_stream_config(0, 0, 0)


def test_multi_gpu_cpp_binary(tmp_path):
    """tests/cpp/test_multi_gpu.cc: the scheduler from caller-side C++ (one std::vector + device list)."""
    exe = tmp_path / "test_multi_gpu"
    cmd = ["g++", "-std=c++14", "-O2", f"-I{ROOT}/include", f"-I{ROOT}/oracle/eigen_standin",
           "-I/usr/local/cuda/include", f"{ROOT}/tests/cpp/test_multi_gpu.cc", "-o", str(exe),
           f"-L{ROOT}/eigencuda_b200/lib", "-leigencuda_b200", f"-Wl,-rpath,{ROOT}/eigencuda_b200/lib"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "ALL PASSED" in res.stdout, res.stdout + res.stderr


# ---------------------------------------------------------------------------------------------
# one large DGEMM, 2D block-partitioned over a device grid (ec_dgemm_2d, BASELINE.json configs[4])
# ---------------------------------------------------------------------------------------------
def _grids():
    g = ec.count_available_gpus()
    out = [([0], 1, 1)]
    if g >= 2:
        out += [([0, 1], 1, 2), ([0, 1], 2, 1), ([1, 0], 0, 0)]
    if g >= 4:
        out += [([0, 1, 2, 3], 2, 2), ([0, 1, 2, 3], 1, 4), ([0, 1, 2, 3], 4, 1)]
    if g >= 8:
        out += [(list(range(8)), 2, 4), (list(range(8)), 4, 2)]
    return out


@pytest.mark.parametrize("shape,panel", [((640, 512, 768), 128), ((515, 389, 1030), 100), ((96, 1200, 64), 0),
                                         ((1, 1, 1), 0), ((300, 200, 7), 3)])
def test_dgemm_2d_matches_oracle_on_every_grid(shape, panel):
    """SUMMA over 1x1, 1x2, 2x1, 2x2, 1x4, 4x1, 2x4, 4x2 (what the box has): results within 1e-12 of
    the long-double oracle, ragged blocks and panels, host upload / download round trip."""
    m, n, k = shape
    A, B = rnd(11, (m, k)), rnd(12, (k, n))
    want = orc.dgemm(A, B, flavour="ld")
    for devices, P, Q in _grids():
        grid = ec.DeviceGrid(devices, P, Q)
        dA, dB, dC = grid.matrix(m, k).upload(A), grid.matrix(k, n).upload(B), grid.matrix(m, n)
        assert np.array_equal(dA.download(), A)                         # blocks round-trip bit-exactly
        l0, b0 = grid.launch_count(), grid.broadcast_count()
        ms = grid.dgemm(dA, dB, dC, panel=panel)
        got = dC.download()
        assert orc.rel_frobenius(got, want) <= TOL, (devices, P, Q)
        assert grid.launch_count() > l0 and ms >= 0.0
        if grid.P * grid.Q > 1:
            assert grid.broadcast_count() > b0                          # the panels went through NCCL
        assert dC.get(m - 1, n - 1) == got[m - 1, n - 1] and dC.get(0, 0) == got[0, 0]
        for x in (dA, dB, dC):
            x.close()
        grid.close()


def test_dgemm_2d_generated_operands_and_shape_errors():
    """Operands synthesised block by block on their owners equal the global generator; the product
    of generated operands matches sampled long-double dot products (how n = 65536 is verified)."""
    n = 1536
    grid = ec.DeviceGrid("all")
    dA, dB, dC = grid.matrix(n, n).fill_uniform(orc.SEED, 0), grid.matrix(n, n).fill_uniform(orc.SEED, 1), grid.matrix(n, n)
    assert np.array_equal(dA.download(), orc.fill_uniform(orc.SEED, 0, (n, n)))
    grid.dgemm(dA, dB, dC, panel=256)
    rng = np.random.default_rng(5)
    for _ in range(12):
        i, j = int(rng.integers(0, n)), int(rng.integers(0, n))
        arow = orc.uniform_strided(orc.SEED, 0, i, n, n).astype(np.longdouble)
        bcol = orc.uniform_strided(orc.SEED, 1, j * n, 1, n).astype(np.longdouble)
        want = float(np.dot(arow, bcol))
        scale = float(np.linalg.norm(arow.astype(np.float64)) * np.linalg.norm(bcol.astype(np.float64)))
        assert abs(dC.get(i, j) - want) <= 1e-13 * scale
    bad = grid.matrix(n + 1, 8)
    before = grid.launch_count()
    with pytest.raises(RuntimeError) as e:
        grid.dgemm(dA, bad, dC)
    assert e.value.code == 1 and grid.launch_count() == before          # before any device work
    with pytest.raises(RuntimeError):
        ec.DeviceGrid([0], 2, 2)
    # host-pointer convenience form
    A, B = rnd(21, (200, 300)), rnd(22, (300, 100))
    assert orc.rel_frobenius(ec.dgemm_2d(A, B), orc.dgemm(A, B, flavour="ld")) <= TOL
    grid.close()


def test_pick_stream_devices_returns_a_measured_subset():
    g = ec.count_available_gpus()
    picked, gbs_picked, gbs_all = ec.pick_stream_devices("all")
    assert 1 <= len(picked) <= g and len(set(picked)) == len(picked) and all(0 <= d < g for d in picked)
    assert gbs_picked >= gbs_all > 5.0
    if len(picked) < g:
        assert gbs_picked > 1.05 * gbs_all                 # a subset is only chosen when clearly faster
    assert ec.pick_stream_devices("all")[0] == picked        # cached: the same answer, no second probe


def test_host_link_probe_reports_a_ceiling():
    got = ec.host_link_probe([0], nbytes=64 << 20, reps=2)
    assert set(got) == {"h2d", "d2h", "bidir_each_way"}
    assert all(5.0 < v < 500.0 for v in got.values()), got        # GB/s: PCIe Gen5 x16 / C2C territory


def test_host_link_probe_on_caller_buffers():
    nbytes = 96 << 20
    a, b = ec.PinnedBuffer((nbytes // 8,)), ec.PinnedBuffer((nbytes // 8,))
    g = ec.count_available_gpus()
    got = ec.host_link_probe_buffers("all", a.ptr, b.ptr, nbytes // g // 8 * 8, 16 << 20, 2)
    assert all(5.0 < v < 1000.0 for v in got.values()), got
    pageable = np.zeros(1 << 20)
    with pytest.raises(RuntimeError):
        ec.host_link_probe_buffers([0], pageable.ctypes.data, b.ptr, 1 << 20)
    a.close()
    b.close()


def test_tensor_stream_rejects_unusable_result_arrays(pipe):
    """caller-supplied outputs are validated, not written through."""
    n = 64
    F, ts = rnd(1, (n, n)), [rnd(2, (n, n)), rnd(3, (n, n))]
    good = [np.empty((n, n), order="F") for _ in ts]
    for bad in ([good[0]],                                             # too few
                [good[0], np.empty((n, n), dtype=np.float32, order="F")],
                [good[0], np.empty((n, n + 1), order="F")],
                [good[0], np.empty((2 * n, n), order="F")[::2]],       # strided
                [good[0], np.empty((n, n + 8), order="C")[:, :n]]):
        with pytest.raises(ValueError):
            pipe.tensor_stream(ec.EC_STREAM_RIGHT, F, ts, results=bad)
    ro = np.empty((n, n), order="F")
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        pipe.tensor_stream(ec.EC_STREAM_RIGHT, F, ts, results=[good[0], ro])
    res = pipe.tensor_stream(ec.EC_STREAM_RIGHT, F, ts, results=good)
    assert res[0] is good[0] and orc.rel_frobenius(good[1], ts[1] @ F) <= TOL


def test_host_pin_registers_caller_memory_in_place(pipe):
    """ec_host_register: matrices the caller already owns, pinned in place, go
    through the direct-DMA paths of copy_to_gpu / conversion / tensor stream; results identical."""
    n, count = 192, 9
    F = orc.fill_uniform(orc.SEED, 0, (n, n))
    ts = [orc.fill_uniform(orc.SEED, i + 1, (n, n)) for i in range(count)]
    outs = [np.empty((n, n), order="F") for _ in range(count)]
    want = orc.tensor_stream("right", F, ts, flavour="fast")
    with ec.host_pin(F, *ts, *outs):
        with ec.host_pin(ts[0]):          # registering twice is not an error
            pass
        res = pipe.tensor_stream(ec.EC_STREAM_RIGHT, F, ts, results=outs)
        for r, w in zip(res, want):
            assert orc.rel_frobenius(r, w) <= TOL
        big = orc.fill_uniform(7, 3, (700, 600))      # 3.4 MB: the chunked download path if pageable
        with ec.host_pin(big):
            d = ec.CudaMatrix(big, pipe.get_stream())
            pipe.synchronize()                        # a pinned source is read asynchronously
            assert np.array_equal(d.to_eigen(), big)
    # unpinned again: the staged paths give the same bytes
    res2 = pipe.tensor_stream(ec.EC_STREAM_RIGHT, F, ts)
    for r, r2 in zip(outs, res2):
        assert np.array_equal(r, r2)


def test_tensor_stream_mixed_host_memory_and_device_fixed(pipe):
    """Pinned inputs with pageable outputs (and the reverse), and a fixed operand that already
    lives on the device (what the multi-GPU broadcast leaves behind): detected per buffer."""
    import ctypes as C
    n, count = 128, 11
    F = orc.fill_uniform(orc.SEED, 0, (n, n))
    dF = ec.CudaMatrix(F, pipe.get_stream())
    pin = ec.PinnedBuffer((n, n, count))
    for i in range(count):
        pin.array[:, :, i] = orc.fill_uniform(orc.SEED, i + 1, (n, n))
    pinned_ts = [pin.array[:, :, i] for i in range(count)]
    pageable_ts = [np.asfortranarray(t.copy()) for t in pinned_ts]
    want = orc.tensor_stream("right", F, pageable_ts, flavour="fast")
    pipe.set_stream_config(3, 3, 2)
    pout = ec.PinnedBuffer((n, n, count))
    for ts, outs in ((pinned_ts, None), (pageable_ts, [pout.array[:, :, i] for i in range(count)])):
        res = pipe.tensor_stream(ec.EC_STREAM_RIGHT, F, ts, results=outs)
        for r, w in zip(res, want):
            assert orc.rel_frobenius(r, w) <= TOL
    # fixed operand passed as a DEVICE pointer
    outs = [np.empty((n, n), order="F") for _ in range(count)]
    ip = (C.c_void_p * count)(*[t.ctypes.data for t in pageable_ts])
    op = (C.c_void_p * count)(*[o.ctypes.data for o in outs])
    pipe.tensor_stream_raw(ec.EC_STREAM_RIGHT, dF.data(), n, n, ip, op, count, n, n)
    for r, w in zip(outs, want):
        assert orc.rel_frobenius(r, w) <= TOL
    pipe.set_stream_config(0, 0, 0)
    pin.close()
    pout.close()


def test_operands_created_on_another_stream_and_two_pipelines(pipe):
    """gemm on pipeline A with matrices created on pipeline B's stream (the reference is ordered by
    the legacy NULL stream there; here the dependency is explicit), and two pipelines used in
    alternation, each with its own stream-K workspace."""
    other = ec.CudaPipeline(0)
    try:
        A, B = rnd(11, (1280, 520)), rnd(12, (520, 1280))
        dA = ec.CudaMatrix(A, other.get_stream())
        dB = ec.CudaMatrix(B, other.get_stream())
        dC = ec.CudaMatrix(1280, 1280, other.get_stream())
        pipe.gemm(dA, dB, dC)                       # runs on pipe's stream, operands from `other`
        got = dC.to_eigen()                         # downloads on `other`'s stream
        want = orc.cpu_product(A, B)
        assert orc.rel_frobenius(got, want) <= TOL
        dC2 = ec.CudaMatrix(1280, 1280, pipe.get_stream())
        for _ in range(4):
            pipe.gemm(dA, dB, dC2)
            other.gemm(dA, dB, dC)
        assert np.array_equal(dC2.to_eigen(), dC.to_eigen())
        assert orc.rel_frobenius(dC.to_eigen(), want) <= TOL
    finally:
        other.close()


def test_tensor_stream_rectangular_reference_order(pipe, ref):
    """T_i (r x c) * F (c x c2): the reference's gemm(cuma_B, cuma_A, cuma_C) order, vs the
    unmodified reference run on the same inputs."""
    r, c, c2, count = 96, 64, 48, 10
    F = orc.fill_uniform(orc.SEED, 0, (c, c2))
    ts = [orc.fill_uniform(orc.SEED, i + 1, (r, c)) for i in range(count)]
    got = pipe.right_multiply(ts, F)
    want, _ = ref.tensor_stream(0, F, ts)
    for g, w in zip(got, want):
        assert g.shape == (r, c2)
        assert orc.rel_frobenius(g, w) <= TOL


def test_tensor_stream_device_resident(pipe):
    n, count = 128, 12
    dF = ec.CudaMatrix(n, n, pipe.get_stream())
    dT = ec.CudaMatrix(n, n * count, pipe.get_stream())
    dR = ec.CudaMatrix(n, n * count, pipe.get_stream())
    pipe.fill_uniform(orc.SEED, 0, dF.data(), n * n, 1)
    pipe.fill_uniform(orc.SEED, 1, dT.data(), n * n, count)
    for kind, name in [(ec.EC_STREAM_RIGHT, "right"), (ec.EC_STREAM_BASIS, "basis")]:
        pipe.tensor_stream_device(kind, dF.data(), n, n, dT.data(), dR.data(), count, n, n)
        got = dR.to_eigen().reshape(n, n, count, order="F")
        F = orc.fill_uniform(orc.SEED, 0, (n, n))
        ts = [orc.fill_uniform(orc.SEED, i + 1, (n, n)) for i in range(count)]
        for i, w in enumerate(orc.tensor_stream(name, F, ts, flavour="fast")):
            assert orc.rel_frobenius(got[:, :, i], w) <= TOL


# ---------------------------------------------------------------------------------------------
# BASELINE.json full sizes, through size-independent properties
# ---------------------------------------------------------------------------------------------
def test_full_size_square_8192_freivalds(pipe):
    """config 2 at n = 8192: C x == A (B x) for random x (O(n^2) check of an O(n^3) product)."""
    n = 8192
    dA, dB = ec.CudaMatrix(n, n, pipe.get_stream()), ec.CudaMatrix(n, n, pipe.get_stream())
    dC = ec.CudaMatrix(n, n, pipe.get_stream())
    pipe.fill_uniform(orc.SEED, 0, dA.data(), n * n, 1)
    pipe.fill_uniform(orc.SEED, 1, dB.data(), n * n, 1)
    pipe.gemm(dA, dB, dC)
    assert pipe.last_kernel().startswith("tma_dmma_NN")
    A, B, Cm = dA.to_eigen(), dB.to_eigen(), dC.to_eigen()
    assert np.array_equal(A[:5, 0], orc.fill_uniform(orc.SEED, 0, (5,)))
    x = np.random.default_rng(0).uniform(-1, 1, (n, 4))
    lhs, rhs = Cm @ x, A @ (B @ x)
    assert np.linalg.norm(lhs - rhs) / np.linalg.norm(rhs) <= TOL
    # sampled elements against long-double dot products
    rng = np.random.default_rng(1)
    for i, j in zip(rng.integers(0, n, 16), rng.integers(0, n, 16)):
        want = float(np.dot(A[i, :].astype(np.longdouble), B[:, j].astype(np.longdouble)))
        assert abs(Cm[i, j] - want) <= 1e-12 * np.linalg.norm(A[i, :]) * np.linalg.norm(B[:, j])


def test_full_size_tensor_stream_4096x512(pipe):
    """config 3: 4096 x (512x512) against a fixed 512x512, resident in HBM; identity, linearity
    and sampled matrices against the oracle."""
    n, count = 512, 4096
    dF = ec.CudaMatrix(n, n, pipe.get_stream())
    dT = ec.CudaMatrix(n, n * count, pipe.get_stream())
    dR = ec.CudaMatrix(n, n * count, pipe.get_stream())
    pipe.fill_uniform(orc.SEED, 0, dF.data(), n * n, 1)
    pipe.fill_uniform(orc.SEED, 1, dT.data(), n * n, count)
    pipe.tensor_stream_device(ec.EC_STREAM_RIGHT, dF.data(), n, n, dT.data(), dR.data(), count, n, n)
    F = orc.fill_uniform(orc.SEED, 0, (n, n))
    one = ec.CudaMatrix(n, n, pipe.get_stream())
    dI = ec.CudaMatrix(np.eye(n, order="F"), pipe.get_stream())

    def pull(i):
        # result i out of the resident tensor through an identity product (itself a property:
        # R_i * I == R_i exactly), without downloading the whole 8 GiB tensor
        pipe.dgemm(0, 0, n, n, n, 1.0, dR.data() + 8 * n * n * i, n, dI.data(), n, 0.0, one.data(), n)
        return one.to_eigen()

    for i in (0, 1, 2047, 4095):
        Ti = orc.fill_uniform(orc.SEED, i + 1, (n, n))
        assert orc.rel_frobenius(pull(i), orc.cpu_product(Ti, F)) <= TOL
    # linearity across the tensor: (T_0 + T_1) F == R_0 + R_1
    R0, R1 = pull(0), pull(1)
    T01 = orc.fill_uniform(orc.SEED, 1, (n, n)) + orc.fill_uniform(orc.SEED, 2, (n, n))
    d01 = ec.CudaMatrix(T01, pipe.get_stream())
    pipe.gemm(d01, dF, one)
    assert orc.rel_frobenius(one.to_eigen(), R0 + R1) <= 4 * TOL


def test_full_size_transposed_1000(pipe):
    """config 4: A^T B and A^T M A at 1000 x 1000 (a few tensor elements), vs the oracle."""
    n, count = 1000, 3
    F = orc.fill_uniform(orc.SEED, 0, (n, n))
    ts = [orc.fill_uniform(orc.SEED, i + 1, (n, n)) for i in range(count)]
    for got, t in zip(pipe.left_multiply(F, ts, transpose=True), ts):
        assert orc.rel_frobenius(got, orc.cpu_product(F.T, t)) <= TOL
    for got, t in zip(pipe.basis_transform(F, ts), ts):
        assert orc.rel_frobenius(got, orc.cpu_product(orc.cpu_product(F.T, t), F)) <= TOL


# ---------------------------------------------------------------------------------------------
# EC_BACKEND_OZAKI_I8: the selectable int8 tensor-core emulation, with its stated (looser) bound
# ---------------------------------------------------------------------------------------------
OZ_TOL = 1e-12   # same bar as the native path on well-scaled data (measured ~5e-15)


@pytest.fixture()
def ozaki(pipe):
    pipe.set_backend(ec.EC_BACKEND_OZAKI_I8)
    yield pipe
    pipe.set_backend(ec.EC_BACKEND_FP64_DMMA)


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("shape", [(256, 192, 320), (200, 136, 520), (130, 70, 1000), (512, 512, 512), (384, 100, 72)])
def test_ozaki_backend_all_layouts(ozaki, ta, tb, shape):
    m, n, k = shape
    got, want = run_dgemm(ozaki, ta, tb, m, n, k, 1.25, -0.5, lda_pad=2, ldb_pad=4, ldc_pad=2, seed=m + k)
    assert ozaki.last_kernel() == "ozaki_i8_s7"
    assert orc.rel_frobenius(got, want) <= OZ_TOL


def test_ozaki_small_shapes_fall_back_to_native(ozaki):
    got, want = run_dgemm(ozaki, False, False, 48, 40, 32, 1.0, 0.0, seed=3)
    assert not ozaki.last_kernel().startswith("ozaki")
    assert orc.rel_frobenius(got, want) <= TOL


def test_ozaki_stated_bound_on_badly_scaled_rows(ozaki):
    """The emulation's bound is norm-wise per row of A and per column of B:
    |E_ij| <= (S+2) * K * 128^-S * 2^Ei * 2^Ej with 2^Ei <= 4 * max_p |a_ip| (S = 7).  Rows whose
    entries span 16 orders of magnitude lose their small entries -- by design -- but stay inside
    that bound; the native path is the one to use when that matters."""
    rng = np.random.default_rng(5)
    m, n, k = 256, 128, 384
    A = np.asfortranarray(rng.uniform(-1, 1, (m, k)) * 10.0 ** rng.integers(-8, 9, (m, k)))
    B = np.asfortranarray(rng.uniform(-1, 1, (k, n)) * 10.0 ** rng.integers(-8, 9, (k, n)))
    dA, dB = ec.CudaMatrix(A, ozaki.get_stream()), ec.CudaMatrix(B, ozaki.get_stream())
    dC = ec.CudaMatrix(m, n, ozaki.get_stream())
    ozaki.gemm(dA, dB, dC)
    assert ozaki.last_kernel() == "ozaki_i8_s7"
    got = dC.to_eigen()
    want = orc.dgemm(A, B, flavour="ld")
    bound = 9 * k * 128.0 ** -7 * np.outer(4 * np.abs(A).max(axis=1), 4 * np.abs(B).max(axis=0))
    assert (np.abs(got - want) <= bound).all()
    # and row/column scaling by powers of two is exact: scaling inputs scales outputs bit for bit
    dA2 = ec.CudaMatrix(A * 2.0 ** 40, ozaki.get_stream())
    ozaki.gemm(dA2, dB, dC)
    assert np.array_equal(dC.to_eigen(), got * 2.0 ** 40)


def test_ozaki_deep_k_is_chunked(ozaki):
    """K beyond 65536 (where 7 planes would overflow the exact int32 accumulators) is cut along k
    and accumulated with beta = 1; a last chunk shorter than 64 still goes through."""
    m, n, k = 128, 64, 65536 + 4096 + 24
    A, B = rnd(3, (m, k)), rnd(4, (k, n))
    Cin = rnd(5, (m, n))
    dA, dB = ec.CudaMatrix(A, ozaki.get_stream()), ec.CudaMatrix(B, ozaki.get_stream())
    dC = ec.CudaMatrix(Cin, ozaki.get_stream())
    ozaki.dgemm(0, 0, m, n, k, 0.5, dA.data(), m, dB.data(), k, -2.0, dC.data(), m)
    assert ozaki.last_kernel() == "ozaki_i8_s7"
    want = orc.dgemm(A, B, alpha=0.5, beta=-2.0, Cin=Cin, flavour="ld")
    assert orc.rel_frobenius(dC.to_eigen(), want) <= OZ_TOL


def test_ozaki_nan_and_zero_rows(ozaki):
    m, n, k = 256, 128, 256
    A, B = rnd(1, (m, k)), rnd(2, (k, n))
    A[7, :] = 0.0
    A[9, 100] = np.nan
    B[3, 5] = np.inf
    dA, dB = ec.CudaMatrix(A, ozaki.get_stream()), ec.CudaMatrix(B, ozaki.get_stream())
    dC = ec.CudaMatrix(m, n, ozaki.get_stream())
    ozaki.gemm(dA, dB, dC)
    got = dC.to_eigen()
    assert np.array_equal(got[7, np.arange(n) != 5], np.zeros(n - 1))     # zero row -> exact zeros
    assert np.isnan(got[9, :]).all() and np.isnan(got[:, 5]).all()       # non-finite input poisons its row / column
    ok = np.ones((m, n), bool); ok[9, :] = False; ok[:, 5] = False
    want = orc.dgemm(np.where(np.isfinite(A), A, 0), np.where(np.isfinite(B), B, 0), flavour="ld")
    assert orc.rel_frobenius(np.where(ok, got, 0), np.where(ok, want, 0)) <= OZ_TOL


def test_ozaki_tensor_stream_and_batch_chunking(ozaki, monkeypatch):
    monkeypatch.setenv("EC_OZAKI_CHUNK", "5")          # 23 matrices -> 5 chunks, ragged tail
    n, count = 192, 23
    F = orc.fill_uniform(orc.SEED, 0, (n, n))
    ts = [orc.fill_uniform(orc.SEED, i + 1, (n, n)) for i in range(count)]
    for kind, name in ((ec.EC_STREAM_RIGHT, "right"), (ec.EC_STREAM_LEFT_T, "left_t"), (ec.EC_STREAM_BASIS, "basis")):
        res = ozaki.tensor_stream(kind, F, ts)
        for r, w in zip(res, orc.tensor_stream(name, F, ts, flavour="fast")):
            assert orc.rel_frobenius(r, w) <= OZ_TOL
    assert ozaki.last_kernel() == "ozaki_i8_s7"


# ---------------------------------------------------------------------------------------------
# seeded fuzz over shapes / layouts / paddings / scalars / batching, both backends
# ---------------------------------------------------------------------------------------------
def _fuzz_case(rng):
    big = rng.random() < 0.3
    hi = 900 if big else 260
    m, n, k = (int(rng.integers(1, hi)) for _ in range(3))
    if rng.random() < 0.5:                       # make TMA-addressable shapes common
        m, n, k = m + m % 2, n + n % 2, k + k % 2
    ta, tb = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    pads = [int(rng.choice([0, 0, 2, 3, 8])) for _ in range(3)]
    alpha = float(rng.choice([1.0, -0.5, 2.25]))
    beta = float(rng.choice([0.0, 0.0, 1.0, -1.5]))
    return m, n, k, ta, tb, pads, alpha, beta


@pytest.mark.parametrize("backend", ["native", "ozaki"])
def test_fuzz_dgemm(pipe, backend):
    rng = np.random.default_rng(20200210 if backend == "native" else 20200211)
    pipe.set_backend(ec.EC_BACKEND_OZAKI_I8 if backend == "ozaki" else ec.EC_BACKEND_FP64_DMMA)
    kernels = set()
    try:
        for case in range(70):
            m, n, k, ta, tb, pads, alpha, beta = _fuzz_case(rng)
            got, want = run_dgemm(pipe, ta, tb, m, n, k, alpha, beta, *pads, seed=1000 + case)
            kernels.add(pipe.last_kernel())
            err = orc.rel_frobenius(got, want)
            assert err <= TOL, (case, m, n, k, ta, tb, pads, alpha, beta, pipe.last_kernel(), err)
    finally:
        pipe.set_backend(ec.EC_BACKEND_FP64_DMMA)
    assert "simt" in kernels and any(x.startswith("tma_dmma") for x in kernels)
    if backend == "ozaki":
        assert "ozaki_i8_s7" in kernels


def test_fuzz_batched(pipe):
    rng = np.random.default_rng(7)
    for case in range(12):
        m, n, k = (int(2 * rng.integers(8, 200)) for _ in range(3))
        batch = int(rng.integers(1, 40))
        ta, tb = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        shared_b = bool(rng.integers(0, 2))
        a_shape = (k, m) if ta else (m, k)
        b_shape = (n, k) if tb else (k, n)
        As = np.asfortranarray(rng.uniform(-1, 1, a_shape + (batch,)))
        Bs = np.asfortranarray(rng.uniform(-1, 1, b_shape + ((1,) if shared_b else (batch,))))
        dA = ec.CudaMatrix(As.reshape(a_shape[0], -1, order="F"), pipe.get_stream())
        dB = ec.CudaMatrix(Bs.reshape(b_shape[0], -1, order="F"), pipe.get_stream())
        dC = ec.CudaMatrix(m, n * batch, pipe.get_stream())
        for backend in (ec.EC_BACKEND_FP64_DMMA, ec.EC_BACKEND_OZAKI_I8):
            pipe.set_backend(backend)
            pipe.dgemm_strided_batched(int(ta), int(tb), m, n, k, 1.0, dA.data(), a_shape[0], a_shape[0] * a_shape[1],
                                       dB.data(), b_shape[0], 0 if shared_b else b_shape[0] * b_shape[1], 0.0,
                                       dC.data(), m, m * n, batch)
            got = dC.to_eigen().reshape(m, n, batch, order="F")
            for i in range(batch):
                want = orc.dgemm(As[:, :, i], Bs[:, :, 0 if shared_b else i], ta, tb, flavour="fast")
                assert orc.rel_frobenius(got[:, :, i], want) <= TOL, (case, m, n, k, batch, ta, tb, shared_b, pipe.last_kernel())
        pipe.set_backend(ec.EC_BACKEND_FP64_DMMA)
