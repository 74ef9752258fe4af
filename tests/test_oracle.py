"""The oracle against the reference's golden vectors (not gpu).

Pins oracle/dgemm_oracle.c before anything is compared with it:
  - the known answers of the reference's tests (src/tests/test_dot.cc:18-26,58-74,86-89);
  - the shape-mismatch throw (src/tests/test_dot.cc:92-103);
  - outputs of the unmodified reference (cuBLAS) captured on a B200 (tests/golden/ref_cublas.npz,
    made by tests/golden/make_ref_golden.py) when that fixture is present;
  - numpy/OpenBLAS as an independent CPU product, and the three accumulation flavours against
    each other, for all op(A)/op(B)/alpha/beta combinations and ragged / empty shapes.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle_py as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _mat(d):
    return np.asfortranarray(np.array(d["row_major"], dtype=np.float64).reshape(d["rows"], d["cols"]))


@pytest.fixture(scope="module")
def readme():
    with open(os.path.join(GOLD, "readme_example.json")) as f:
        return json.load(f)


def test_readme_column_major_bytes(readme):
    # the comma initialiser fills row-major, storage is column-major
    assert _mat(readme["A"]).ravel(order="K").tolist() == readme["column_major_bytes"]["A"]
    assert _mat(readme["tensor"][0]).ravel(order="K").tolist() == readme["column_major_bytes"]["B"]
    assert _mat(readme["expected"][0]).ravel(order="K").tolist() == readme["column_major_bytes"]["X"]


@pytest.mark.parametrize("flavour", ["plain", "ld", "fast"])
def test_readme_known_answers(readme, flavour):
    A = _mat(readme["A"])
    for t, x in zip(readme["tensor"], readme["expected"]):
        got = orc.dgemm(_mat(t), A, flavour=flavour)       # gemm(cuma_B, cuma_A, cuma_C)
        assert np.array_equal(got, _mat(x))                # small integers: exact
        assert orc.is_approx(got, _mat(x))


@pytest.mark.parametrize("flavour", ["plain", "ld", "fast"])
def test_readme_streaming_loop(readme, flavour):
    A = _mat(readme["A"])
    res = orc.tensor_stream("right", A, [_mat(t) for t in readme["tensor"]], flavour=flavour)
    for r, x in zip(res, readme["expected"]):
        assert np.array_equal(r, _mat(x))


def test_wrong_shape_raises(readme):
    a, b = readme["wrong_shape"]["A"], readme["wrong_shape"]["B"]
    with pytest.raises(orc.ShapeMismatch):
        orc.dgemm(np.ones(a), np.ones(b))
    with pytest.raises(orc.ShapeMismatch):
        orc.tensor_stream("right", np.ones((5, 5)), [np.ones((3, 2))])


def test_is_approx_is_eigens_formula():
    a = np.asfortranarray(np.arange(1.0, 13.0).reshape(3, 4))
    b = a * (1 + 5e-13)
    assert orc.is_approx(a, b)                       # 5e-13 relative: inside 1e-12
    assert not orc.is_approx(a, a * (1 + 5e-12))
    d = np.linalg.norm(a - b) / min(np.linalg.norm(a), np.linalg.norm(b))
    assert abs(orc.rel_frobenius(a, b) - d) < 1e-18 + 1e-6 * d


@pytest.mark.parametrize("transa", [False, True])
@pytest.mark.parametrize("transb", [False, True])
@pytest.mark.parametrize("shape", [(200, 200, 200), (3, 2, 2), (65, 130, 31), (1, 7, 1), (17, 1, 9)])
def test_flavours_agree_with_numpy(transa, transb, shape):
    m, n, k = shape
    rng = np.random.default_rng(m * 1000 + n * 10 + k)
    A = np.asfortranarray(rng.uniform(-1, 1, (k, m) if transa else (m, k)))
    B = np.asfortranarray(rng.uniform(-1, 1, (n, k) if transb else (k, n)))
    Cin = np.asfortranarray(rng.uniform(-1, 1, (m, n)))
    want = 0.75 * (A.T if transa else A) @ (B.T if transb else B) - 1.25 * Cin
    for fl in ("plain", "ld", "fast"):
        got = orc.dgemm(A, B, transa, transb, alpha=0.75, beta=-1.25, Cin=Cin, flavour=fl)
        assert orc.rel_frobenius(got, want) <= 1e-13, fl


def test_beta_zero_does_not_read_c():
    A, B = np.ones((4, 3), order="F"), np.ones((3, 5), order="F")
    poisoned = np.full((4, 5), np.nan, order="F")
    for fl in ("plain", "ld", "fast"):
        got = orc.dgemm(A, B, beta=0.0, Cin=poisoned, flavour=fl)
        assert np.array_equal(got, np.full((4, 5), 3.0))


def test_empty_shapes():
    assert orc.dgemm(np.zeros((0, 3)), np.zeros((3, 4))).shape == (0, 4)
    assert np.array_equal(orc.dgemm(np.zeros((2, 0)), np.zeros((0, 4))), np.zeros((2, 4)))
    assert orc.tensor_stream("right", np.eye(2), []) == []


@pytest.mark.parametrize("kind", ["right", "left", "left_t", "basis"])
def test_tensor_stream_kinds(kind):
    rng = np.random.default_rng(11)
    F = np.asfortranarray(rng.uniform(-1, 1, (24, 24)))
    ts = [np.asfortranarray(rng.uniform(-1, 1, (24, 24))) for _ in range(5)]
    want = {"right": lambda t: t @ F, "left": lambda t: F @ t, "left_t": lambda t: F.T @ t,
            "basis": lambda t: F.T @ t @ F}[kind]
    for fl in ("plain", "ld", "fast"):
        for t, r in zip(ts, orc.tensor_stream(kind, F, ts, flavour=fl)):
            assert orc.rel_frobenius(r, want(t)) <= 1e-13


def test_fill_uniform_is_deterministic_and_uniform():
    a = orc.fill_uniform(orc.SEED, 3, (64, 32))
    b = orc.fill_uniform(orc.SEED, 3, (64, 32))
    c = orc.fill_uniform(orc.SEED, 4, (64, 32))
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.min() >= -1.0 and a.max() < 1.0 and abs(a.mean()) < 0.05
    # element i depends only on (seed, id, column-major index), not on the shape
    assert np.array_equal(a.ravel(order="F")[:100], orc.fill_uniform(orc.SEED, 3, (100, 1)).ravel())
    # known answers of the generator itself (guards the recipe the CUDA generator must match)
    gold = json.load(open(os.path.join(GOLD, "uniform_kat.json")))
    for e in gold["values"]:
        assert orc.lib().oracle_uniform_at(gold["seed"], e["matrix_id"], e["index"]) == float.fromhex(e["hex"])


def test_against_reference_cublas_outputs():
    """Oracle vs the UNMODIFIED reference run on a B200 (fixture committed from a gpurun call)."""
    path = os.path.join(GOLD, "ref_cublas.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ref_cublas.npz not generated yet (needs a GPU box)")
    from tests.golden.make_ref_golden import CASES, STREAM
    z = np.load(path)
    for cid, (name, m, k, n) in enumerate(CASES):
        A = orc.fill_uniform(orc.SEED, 1000 + 2 * cid, (m, k))
        B = orc.fill_uniform(orc.SEED, 1001 + 2 * cid, (k, n))
        for fl in ("plain", "ld", "fast"):
            assert orc.rel_frobenius(orc.dgemm(A, B, flavour=fl), z[name]) <= 1e-12, (name, fl)
    _, cnt, r, c = STREAM
    F = orc.fill_uniform(orc.SEED, 0, (c, c))
    ts = [orc.fill_uniform(orc.SEED, i + 1, (r, c)) for i in range(cnt)]
    for got, want in zip(orc.tensor_stream("right", F, ts), z["stream_right"]):
        assert orc.rel_frobenius(got, want) <= 1e-12
    for got, want in zip(orc.tensor_stream("left", F, ts), z["stream_left"]):
        assert orc.rel_frobenius(got, want) <= 1e-12
    assert np.array_equal(z["roundtrip"], ts[0])
