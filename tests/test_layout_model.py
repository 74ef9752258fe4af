"""CPU model of the hot kernel's shared-memory index algebra (not gpu).

eigencuda_b200/csrc/dgemm_kernels.cuh stages operand tiles with 128B-swizzled TMA boxes and reads
DMMA fragments with permuted rows / interleaved k so that every LDS.128 is bank-conflict-free.
This test restates those formulas in numpy -- TMA swizzle model, fragment addresses, the
mma.m8n8k4 lane contract, the epilogue's (row, col) mapping -- and checks, for all four
op(A)/op(B) layouts, that (a) the model reproduces A @ B for a full 128x128x16 stage and (b) no
quarter-warp of any LDS.128 touches the same 16-byte bank group twice.  It guards the algebra;
the CUDA kernel itself is checked against the oracle in test_gpu_parity.py.
"""
import numpy as np
import pytest

BM = 128
BK = 16


def swz(off):
    """128B swizzle: byte-address bits [4:6] ^= bits [7:9] (tile base is 1024-aligned)."""
    return off ^ (((off >> 7) & 7) << 4)


def tma_fill(tile, kmajor):
    """tile[mn, k] (128 x 16) -> 16 KiB smem image as float64[2048], as the TMA boxes lay it out."""
    smem = np.full(2048, np.nan)
    for mn in range(BM):
        for k in range(BK):
            if kmajor:      # one box {16 k, 128 mn}: row mn = 128 B
                off = mn * 128 + k * 8
            else:           # eight boxes {16 mn, 16 k}: sub-tile mn>>4, row k = 128 B
                off = (mn >> 4) * 2048 + k * 128 + (mn & 15) * 8
            smem[swz(off) // 8] = tile[mn, k]
    assert not np.isnan(smem).any()
    return smem


def frag_mn(kmajor, blk, g):
    if kmajor:
        return 8 * blk + (g >> 1) + 4 * (g & 1)
    return 16 * (blk >> 1) + 2 * g + (blk & 1)


def load_frags(smem, kmajor, w0, kk, nblk, conflicts):
    """Returns fe[blk][lane], fo[blk][lane] exactly as load_frags<> in the .cuh computes them."""
    fe = np.zeros((nblk, 32))
    fo = np.zeros((nblk, 32))
    def lds128(addrs):  # addrs[lane] byte address; check the quarter-warp bank rule
        for q in range(4):
            groups = [(addrs[l] >> 4) & 7 for l in range(8 * q, 8 * q + 8)]
            if len(set(groups)) != 8:
                conflicts.append((kmajor, w0, kk, q, groups))
        return [(smem[a // 8], smem[a // 8 + 1]) for a in addrs]
    if kmajor:
        for b in range(nblk):
            addrs = []
            for lane in range(32):
                g, c = lane >> 2, lane & 3
                pg = (g >> 1) + 4 * (g & 1)
                addrs.append((w0 + pg) * 128 + ((((kk >> 1) + c) ^ pg) << 4) + b * 1024)
            v = lds128(addrs)
            for lane in range(32):
                fe[b, lane], fo[b, lane] = v[lane]
    else:
        for jb in range(nblk // 2):
            ae, ao = [], []
            for lane in range(32):
                g, c = lane >> 2, lane & 3
                ke, ko = kk + 2 * c, kk + 2 * c + 1
                ae.append((w0 >> 4) * 2048 + ke * 128 + ((g ^ (ke & 7)) << 4) + jb * 2048)
                ao.append((w0 >> 4) * 2048 + ko * 128 + ((g ^ (ko & 7)) << 4) + jb * 2048)
            ve, vo = lds128(ae), lds128(ao)
            for lane in range(32):
                fe[2 * jb, lane], fe[2 * jb + 1, lane] = ve[lane]
                fo[2 * jb, lane], fo[2 * jb + 1, lane] = vo[lane]
    return fe, fo


def dmma(acc, a, b):
    """mma.m8n8k4: lane 4g+c holds A[g][c], B[c][g]; D[g][2c], D[g][2c+1] live in lane 4g+c."""
    A = a.reshape(8, 4)            # [g][c]
    B = b.reshape(8, 4).T          # [c][j]
    acc += A @ B                   # acc[g][j]


@pytest.mark.parametrize("a_kmajor", [False, True])
@pytest.mark.parametrize("b_kmajor", [False, True])
def test_stage_product_and_bank_conflicts(a_kmajor, b_kmajor):
    rng = np.random.default_rng(7)
    At = rng.standard_normal((BM, BK))     # op(A) tile, [m, k]
    Bt = rng.standard_normal((BM, BK))     # op(B) tile, [n, k]
    sA, sB = tma_fill(At, a_kmajor), tma_fill(Bt, b_kmajor)
    Cref = At @ Bt.T
    C = np.full((BM, BM), np.nan)
    conflicts = []
    for warp in range(8):
        wm0, wn0 = (warp & 1) * 64, (warp >> 1) * 32
        acc = np.zeros((8, 4, 8, 8))       # [i][j][g][col-in-block]
        for kk in (0, 8):
            ae, ao = load_frags(sA, a_kmajor, wm0, kk, 8, conflicts)
            be, bo = load_frags(sB, b_kmajor, wn0, kk, 4, conflicts)
            for i in range(8):
                for j in range(4):
                    dmma(acc[i, j], ae[i], be[j])
                    dmma(acc[i, j], ao[i], bo[j])
        for i in range(8):
            for j in range(4):
                for g in range(8):
                    for q in range(8):     # q = 2c+e, the D-fragment column index
                        row = wm0 + frag_mn(a_kmajor, i, g)
                        col = wn0 + frag_mn(b_kmajor, j, q)
                        assert np.isnan(C[row, col]), "two fragments map to one C element"
                        C[row, col] = acc[i, j, g, q]
    assert not conflicts, conflicts[:3]
    assert not np.isnan(C).any()
    np.testing.assert_allclose(C, Cref, rtol=1e-13, atol=1e-13)
