"""numpy restatement of the arithmetic of the EC_BACKEND_OZAKI_I8 path (eigencuda_b200/csrc/ozaki_kernels.cuh:
scale_from_amax, ozaki_split_kernel, the digit-plane products and the epilogue's recombination), checked on
the CPU against the long-double oracle: the error bound DESIGN.md section 3.3 states for the backend,

    |E_ij| <= (S + 2) * K * 128^-S * 2^Ei * 2^Ej        (2^Ei, 2^Ej: the row / column scales),

is a property of the SCHEME, so it can be proved without a GPU; the GPU suite then checks that the kernels
implement the scheme (tests/test_gpu_parity.py::test_ozaki_*).  Also pinned here: digits fit int8, the
remainder recursion is exact, and the int32 accumulators cannot overflow for K <= 65536.
"""
import numpy as np
import pytest

from oracle import oracle_py as orc

S = 7


def row_scales(X):
    """2^E per row with |x| * 2^-E < 1/2 for every entry (E = exponent of the row's abs-max + 1; 1 for a zero row)."""
    amax = np.abs(X).max(axis=1)
    _, e = np.frexp(amax)                        # amax = m * 2^e, m in [1/2, 1)
    return np.where(amax == 0, 1.0, np.ldexp(1.0, np.clip(e + 1, -1000, 1000)))


def split(X, scales):
    """S balanced base-128 digit planes of X / scale: x -> 128 x -> rint (ties to even, as the 1.5 * 2^52
    trick rounds) -> exact remainder."""
    x = X / scales[:, None]                      # exact: the scale is a power of two
    assert (np.abs(x) < 0.5).all()
    planes = []
    for _ in range(S):
        y = x * 128.0                            # exact
        d = np.rint(y)
        x = y - d                                # exact, |x| <= 1/2
        assert (np.abs(d) <= 64).all() and (np.abs(x) <= 0.5).all()
        planes.append(d.astype(np.int8))
    return planes, x                             # x: what is left, |.| <= 1/2 in units of 128^-S


def emulated_product(A, B):
    """C ~= A * B through the digit planes; also returns the scales."""
    sa, sb = row_scales(A), row_scales(B.T)
    pa, ra = split(A, sa)
    pb, rb = split(B.T, sb)
    # the split is exact up to the stated truncation: x = sum_s d_s 128^-(s+1) + r * 128^-S
    rec = sum(p.astype(np.float64) * 128.0 ** -(s + 1) for s, p in enumerate(pa)) + ra * 128.0 ** -S
    assert np.array_equal(rec * sa[:, None], A)
    acc = np.zeros(A.shape[:1] + B.shape[1:], dtype=np.float64)
    wgt = 1.0 / 16384.0                          # 128^-(d+2)
    for d in range(S):                           # one TMEM accumulator per weight d = s + t; s + t >= S dropped
        acc_d = sum(pa[s].astype(np.int64) @ pb[d - s].astype(np.int64).T for s in range(d + 1))
        assert np.abs(acc_d).max() < 2 ** 31     # exact in int32
        acc = acc_d.astype(np.float64) * wgt + acc   # the product is exact: this is the epilogue's fma
        wgt /= 128.0
    return (sa[:, None] * sb[None, :]) * acc, sa, sb


def check_bound(A, B):
    got, sa, sb = emulated_product(A, B)
    want = orc.dgemm(np.asfortranarray(A), np.asfortranarray(B), flavour="ld")
    K = A.shape[1]
    bound = (S + 2) * K * 128.0 ** -S * sa[:, None] * sb[None, :]
    err = np.abs(got - want)
    assert (err <= bound).all(), float((err / bound).max())
    return float((err / bound).max()), orc.rel_frobenius(got, want)


@pytest.mark.parametrize("m,n,k", [(64, 48, 32), (33, 65, 512), (128, 64, 1000)])
def test_bound_on_well_scaled_data(m, n, k):
    rng = np.random.default_rng(m + n + k)
    worst, rel = check_bound(rng.uniform(-1, 1, (m, k)), rng.uniform(-1, 1, (k, n)))
    assert rel <= 1e-13                          # well-scaled data: close to native FP64 (measured 5e-15 on the GPU)


def test_bound_on_badly_scaled_rows_and_columns():
    """entries x 10^U[-8, 8].  Small entries of a row lose digits -- by design -- but the
    result stays inside the row/column bound."""
    rng = np.random.default_rng(5)
    m, n, k = 96, 80, 384
    A = rng.uniform(-1, 1, (m, k)) * 10.0 ** rng.integers(-8, 9, (m, k))
    B = rng.uniform(-1, 1, (k, n)) * 10.0 ** rng.integers(-8, 9, (k, n))
    check_bound(A, B)


def test_bound_on_adversarial_digits():
    """Entries just below half the scale, same sign everywhere: no cancellation, every truncation and every
    dropped digit pair pushes the same way."""
    m, n, k = 32, 32, 777
    A = np.full((m, k), 0.49999999999, dtype=np.float64)
    B = np.full((k, n), 0.3333333333333333, dtype=np.float64)
    worst, _ = check_bound(A, B)
    assert worst > 1e-3                          # the bound is not vacuous: this case uses a visible part of it
    check_bound(-A, B)


def test_zero_rows_and_powers_of_two():
    rng = np.random.default_rng(1)
    A = rng.uniform(-1, 1, (16, 64))
    A[3, :] = 0.0
    A[5, :] = np.ldexp(1.0, rng.integers(-30, 30, 64))   # exact powers of two
    B = rng.uniform(-1, 1, (64, 8))
    got, _, _ = emulated_product(A, B)
    assert np.array_equal(got[3], np.zeros(8))
    check_bound(A, B)


def test_int32_accumulators_cannot_overflow_up_to_the_chunk_length():
    # worst case per accumulator: S digit pairs share weight d = S - 1, |d * d| <= 64 * 64 = 4096 (the header's
    # 4225 = 65^2 is the looser figure it budgets with), K <= 65536 per launch (deeper products are chunked)
    assert S * 65536 * 4225 < 2 ** 31
