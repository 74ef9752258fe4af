"""Host-side logic that needs no GPU: shapes, sharding, the 2-rank gloo path (not gpu)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import eigencuda_b200 as ec
from eigencuda_b200 import shard
from oracle import oracle_py as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for count in (0, 1, 3, 7, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(count, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 4, 4)


def test_result_shape_and_mismatch():
    assert ec.result_shape(ec.EC_STREAM_RIGHT, (2, 2), (3, 2)) == (3, 2)      # README example
    assert ec.result_shape(ec.EC_STREAM_LEFT, (4, 3), (3, 5)) == (4, 5)
    assert ec.result_shape(ec.EC_STREAM_LEFT_T, (3, 4), (3, 5)) == (4, 5)
    assert ec.result_shape(ec.EC_STREAM_BASIS, (6, 4), (6, 6)) == (4, 4)
    with pytest.raises(ec.EigenCudaError):
        ec.result_shape(ec.EC_STREAM_RIGHT, (5, 5), (3, 2))                   # test_dot.cc:92-103
    with pytest.raises(RuntimeError):   # EigenCudaError is a RuntimeError, like std::runtime_error
        ec.result_shape(ec.EC_STREAM_BASIS, (6, 4), (6, 5))


def test_pipeline_is_not_copyable():
    import copy
    p = object.__new__(ec.CudaPipeline)
    with pytest.raises(TypeError):
        copy.copy(p)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    from eigencuda_b200 import shard
    from oracle import oracle_py as orc
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    count, n = 11, 12
    # rank 0 owns the fixed operand; everyone else starts with garbage and receives it
    F = torch.from_numpy(orc.fill_uniform(orc.SEED, 0, (n, n)).copy()) if rank == 0 else torch.full((n, n), float("nan"), dtype=torch.float64)
    shard.broadcast_fixed(F, src=0)
    lo, hi = shard.shard_range(count, rank, world)
    ts = [orc.fill_uniform(orc.SEED, i + 1, (n, n)) for i in range(lo, hi)]
    # the sharded job itself (the oracle stands in for the GPU product in this CPU test)
    res = orc.tensor_stream("right", F.numpy(), ts)
    counts = shard.gather_counts(len(res))
    assert sum(counts) == count and counts[rank] == hi - lo, counts
    chk = torch.tensor([sum(float(r.sum()) for r in res)], dtype=torch.float64)
    dist.all_reduce(chk)
    if rank == 0:
        Fh = orc.fill_uniform(orc.SEED, 0, (n, n))
        want = sum(float((orc.fill_uniform(orc.SEED, i + 1, (n, n)) @ Fh).sum()) for i in range(count))
        assert abs(chk.item() - want) < 1e-9, (chk.item(), want)
        print("GLOO_OK", counts)
    dist.destroy_process_group()
""")


def test_two_rank_gloo_shard_and_broadcast(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "GLOO_OK [6, 5]" in res.stdout or "GLOO_OK [5, 6]" in res.stdout, res.stdout


SUMMA_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    from eigencuda_b200 import summa
    from oracle import oracle_py as orc
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    m, n, k = 45, 38, 52                      # ragged on purpose: nothing divides evenly
    L = summa.Layout(m, n, k, world, rank)
    rg, cg = summa.make_groups(L)
    A = orc.fill_uniform(orc.SEED, 0, (m, k)); B = orc.fill_uniform(orc.SEED, 1, (k, n))
    # local blocks as (cols, rows) row-major views == column-major storage
    A_loc = torch.from_numpy(np.ascontiguousarray(A[L.m0:L.m1, L.ka0:L.ka1].T))
    B_loc = torch.from_numpy(np.ascontiguousarray(B[L.kb0:L.kb1, L.n0:L.n1].T))
    C_loc = torch.full((L.c_shape[1], L.c_shape[0]), float("nan"), dtype=torch.float64)

    def local_gemm(mm, nn, kk, At, lda, Bt, ldb, beta, Ct, ldc):
        # the CPU oracle stands in for the GPU kernel in this host-logic test
        Am = At.numpy().reshape(-1)[: lda * kk].reshape(kk, lda).T[:mm]
        Bm = Bt.numpy().reshape(-1)[: ldb * nn].reshape(nn, ldb).T[:kk]
        Cm = Ct.numpy().T
        Cm[:] = orc.dgemm(Am, Bm, beta=beta, Cin=(Cm if beta != 0 else None))

    summa.summa(L, A_loc, B_loc, C_loc, local_gemm, kb=7, row_group=rg, col_group=cg)
    want = (A @ B)[L.m0:L.m1, L.n0:L.n1]
    err = orc.rel_frobenius(C_loc.numpy().T, want)
    assert err <= 1e-13, (rank, err)
    ok = torch.tensor([1.0])
    dist.all_reduce(ok)
    if rank == 0:
        print("SUMMA_OK", world, (L.P, L.Q), int(ok.item()))
    dist.destroy_process_group()
""")


@pytest.mark.parametrize("world", [2, 4])
def test_summa_partition_gloo(tmp_path, world):
    """The 2D block partition + panel broadcast logic (BASELINE.json configs[4]) on CPU ranks."""
    script = tmp_path / "summa_worker.py"
    script.write_text(SUMMA_WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=280, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stdout + res.stderr
    assert f"SUMMA_OK {world}" in res.stdout, res.stdout


def test_summa_layout_covers_everything():
    from eigencuda_b200 import summa
    for world in (1, 2, 4, 8):
        m, n, k = 1000, 777, 555
        seenC = np.zeros((m, n), dtype=int)
        for r in range(world):
            L = summa.Layout(m, n, k, world, r)
            seenC[L.m0:L.m1, L.n0:L.n1] += 1
            pan = L.panels(64)
            assert pan[0][0] == 0 and pan[-1][1] == k and all(a[1] == b[0] for a, b in zip(pan, pan[1:]))
            for k0, k1 in pan:       # a panel never straddles an ownership boundary
                q, lo = L.a_owner_col(k0); assert lo <= k0 and k1 <= summa.block_range(k, L.Q, q)[1]
                p, lo = L.b_owner_row(k0); assert lo <= k0 and k1 <= summa.block_range(k, L.P, p)[1]
        assert (seenC == 1).all()


# ---------------------------------------------------------------------------------------------
# host-side copy pool (csrc/copy_pool.hpp): plain C++, stress-tested on the CPU
# ---------------------------------------------------------------------------------------------
def test_copy_pool_stress(tmp_path):
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_copy_pool")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(root, "tests", "cpp", "test_copy_pool.cc"),
                    "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0 and "ALL PASSED" in out.stdout, out.stdout + out.stderr


# ---------------------------------------------------------------------------------------------
# bench.py contract that can be checked without a GPU: ONE JSON line on stdout
# ---------------------------------------------------------------------------------------------
def test_bench_reference_arm_prints_exactly_one_json_line():
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["value"] > 0
    for key in ("metric", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "config", "cpu_baseline", "e2e"):
        assert key in line
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # under torchrun only rank 0 runs it; the other ranks exit 0 without output
    env.update({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


# ---------------------------------------------------------------------------------------------
# the dispatcher's cost model, queried without a device: the choices that were measured best on
# a B200 (profiles/r01_small_probe.txt, r01_batched_small_probe.txt, r01_skinny_probe.txt) stay put
# ---------------------------------------------------------------------------------------------
L, S, XS, T, N16 = 1, 2, 3, 4, 5


@pytest.mark.parametrize("shape,expected", [
    ((512, 512, 512, 4096), (L, False)),        # the bench workload: whole 128x128 tiles
    ((8192, 8192, 8192, 1), (L, True)),         # big square: stream-K evens out the last wave
    ((2048, 2048, 2048, 1), (L, True)),
    ((1024, 1024, 1024, 1), (S, True)),
    ((768, 768, 768, 1), (XS, False)),
    ((512, 512, 512, 1), (T, False)),           # small squares: only the 32x32 tile fills 148 SMs
    ((256, 256, 256, 1), (T, False)),
    ((32, 32, 32, 131072), (T, False)),         # tensors of tiny matrices
    ((16, 16, 16, 524288), (T, False)),
    ((64, 64, 64, 32768), (S, False)),
    ((96, 96, 96, 14563), (T, False)),
    ((128, 128, 128, 8192), (L, False)),
    ((64, 64, 1 << 20, 1), (S, True)),          # inner-product shape: stream-K over the whole grid
    ((1 << 20, 32, 32, 1), (XS, False)),        # tall and skinny: HBM-bound
    ((1 << 20, 16, 128, 1), (N16, False)),      # at most 16 columns: the 128x16 tile (no padded DMMAs)
    ((1 << 21, 8, 64, 1), (N16, False)),
    ((1 << 20, 128, 16, 1), (L, False)),
])
def test_cost_model_choices(shape, expected, monkeypatch):
    monkeypatch.delenv("EC_FORCE_TILE", raising=False)
    monkeypatch.delenv("EC_STREAM_K", raising=False)
    assert ec.plan_dgemm(*shape) == expected


def test_cost_model_overrides(monkeypatch):
    monkeypatch.setenv("EC_FORCE_TILE", "2")
    assert ec.plan_dgemm(8192, 8192, 8192)[0] == S
    monkeypatch.setenv("EC_STREAM_K", "0")
    assert ec.plan_dgemm(8192, 8192, 8192) == (S, False)
    with pytest.raises(ec.EigenCudaError):
        ec.plan_dgemm(0, 8, 8)


def test_multi_gpu_shard_plan_covers_every_index_once():
    """ec_plan_shard: the ranges the multi-GPU scheduler hands to its devices (same rule as
    shard.shard_range) are contiguous, ordered, balanced to within one and cover the tensor."""
    import eigencuda_b200 as ec
    from eigencuda_b200 import shard
    for count in (0, 1, 3, 7, 8, 37, 4096, 4097):
        for ndev in (1, 2, 3, 4, 8):
            ranges = [ec.plan_shard(count, ndev, i) for i in range(ndev)]
            assert ranges == [shard.shard_range(count, i, ndev) for i in range(ndev)]
            assert ranges[0][0] == 0 and ranges[-1][1] == count
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ec.EigenCudaError):
        ec.plan_shard(4, 2, 2)


def test_tensor_stream_result_validation_needs_no_gpu():
    """The checks on caller-supplied result arrays run before the library is called."""
    import numpy as np
    import eigencuda_b200 as ec
    pipe = ec.CudaPipeline.__new__(ec.CudaPipeline)      # no device: validation must come first
    pipe._h = None
    n = 8
    F, ts = np.zeros((n, n), order="F"), [np.zeros((n, n), order="F")]
    for bad in ([], [np.zeros((n, n), dtype=np.float32, order="F")], [np.zeros((n, n), order="C")],
                [np.zeros((n + 1, n), order="F")], [[0.0] * n]):
        with pytest.raises(ValueError):
            pipe.tensor_stream(ec.EC_STREAM_RIGHT, F, ts, results=bad)
    with pytest.raises(ValueError):
        pipe.tensor_stream(ec.EC_STREAM_RIGHT, F, ts, devices=[])


def test_chunk_plan_covers_the_tensor_with_short_ends():
    """ec_plan_chunks: boundaries strictly increase from 0 to count, no chunk exceeds the ring slot,
    and the default plan starts and ends with chunks an eighth of the slot (the transfers that
    nothing can overlap)."""
    import eigencuda_b200 as ec
    for count in (0, 1, 5, 31, 32, 33, 191, 192, 193, 1000, 4096, 4097):
        for chunk in (1, 4, 7, 8, 32, 33, 100):
            for ramp in (False, True):
                starts = ec.plan_chunks(count, chunk, ramp)
                assert starts[0] == 0 and starts[-1] == count
                sizes = [b - a for a, b in zip(starts, starts[1:])]
                assert all(0 < sz <= chunk for sz in sizes), (count, chunk, ramp, sizes)
                if not ramp:
                    assert all(sz == chunk for sz in sizes[:-1])
                elif chunk >= 8 and count >= 6 * chunk:
                    assert sizes[0] == chunk // 8 and sizes[1] == chunk // 4 and sizes[2] == chunk // 2
                    assert sizes[-1] <= max(chunk // 8, 1) + 0 or sizes[-1] <= chunk // 2
                    assert max(sizes) == chunk
                    assert len(sizes) <= count // chunk + 7


def test_fastdiv_matches_integer_division():
    """The kernels locate a tile (problem, tile row, tile column) with a multiply-and-shift division by an
    invariant divisor (FastDiv, csrc/dgemm_kernels.cuh; ec_debug_fastdiv runs the same routine on the host):
    exact for every dividend below 2^31 and every divisor in [1, 2^31)."""
    import random
    from eigencuda_b200 import _lib
    lib = _lib.load()
    rng = random.Random(7)
    top = (1 << 31) - 1
    divisors = [1, 2, 3, 5, 7, 16, 17, 148, 641, 4096, 65535, 65536, 65537, (1 << 30) - 1, 1 << 30, (1 << 30) + 1, top]
    divisors += [rng.randrange(1, 1 << rng.randrange(1, 32)) for _ in range(300)]
    for d in divisors:
        d = max(1, min(d, top))
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, top, top - 1, (top // d) * d, max(0, (top // d) * d - 1)]
        ns += [rng.randrange(0, 1 << 31) for _ in range(40)]
        for n in ns:
            n = min(n, top)
            assert lib.ec_debug_fastdiv(n, d) == n // d, (n, d)


def test_panel_plan_of_the_2d_dgemm_respects_ownership_on_every_grid():
    """ec_plan_panels (csrc/dgemm_2d.inl: summa_panels): on every P x Q grid up to 4 x 8 -- more shapes than any box
    offers -- the panels tile [0, k) exactly, none is wider than asked, and none straddles a boundary between two
    column blocks of A (Q parts) or two row blocks of B (P parts): each panel has ONE owner column and ONE owner
    row, which is what lets its broadcasts start from a single root."""
    import eigencuda_b200 as ec
    for P in (1, 2, 3, 4):
        for Q in (1, 2, 3, 4, 8):
            for k in (0, 1, 7, 100, 4096, 4097, 12289, 65536):
                for panel in (0, 1, 64, 1000, 4096):
                    starts = ec.plan_panels(k, P, Q, panel)
                    width = panel or 4096
                    assert starts[0] == 0 and starts[-1] == k
                    assert all(0 < b - a <= width for a, b in zip(starts, starts[1:])) or k == 0
                    cuts = {ec.plan_shard(k, Q, q)[0] for q in range(Q)} | {ec.plan_shard(k, P, p)[0] for p in range(P)}
                    for a, b in zip(starts, starts[1:]):
                        assert not any(a < c < b for c in cuts), (P, Q, k, panel, a, b)
    # the headline case: n = 65536 on 2 x 4, panels of 4096 -> 16 panels, 4 per owner column, 8 per owner row
    assert ec.plan_panels(65536, 2, 4, 4096) == list(range(0, 65536 + 1, 4096))
    from eigencuda_b200._lib import lib
    assert lib.ec_plan_panels(-1, 1, 1, 0, None, 0) == -1 and lib.ec_plan_panels(8, 0, 1, 0, None, 0) == -1
