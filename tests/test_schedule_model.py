"""Integer model of the stream-K schedule and its fix-up (eigencuda_b200/csrc/dgemm_kernels.cuh:
`Schedule`, "stream-K fix-up"), checked exhaustively on the CPU for many (tiles, k-blocks, grid)
combinations.  The device code uses exactly these formulas; what is proved here:

  * the runs [S*d/G, S*(d+1)/G) of the CTAs cut the region into segments that cover every
    k-iteration of every tile exactly once;
  * `owner` and `last` computed from the tile alone (as every CTA does, without communication)
    are the first and the last rank whose run meets the tile;
  * in the two-level sum every contributor's partial is added exactly once, waits only ever point
    to higher ranks (so they cannot form a cycle), and the order of additions is a function of
    the ranks alone (bit-reproducible).
"""
import random

import pytest

SK_TREE_MIN, SK_TREE_GROUP = 8, 16      # dgemm_kernels.cuh


def segments(rank, G, sk_tiles, kblocks):
    """Schedule::next for the stream-K region: list of (tile, kb0, kb1)."""
    S = sk_tiles * kblocks
    it, end = S * rank // G, S * (rank + 1) // G
    out = []
    while it < end:
        tile = it // kblocks
        kb0 = it - tile * kblocks
        kb1 = min(kblocks, kb0 + (end - it))
        out.append((tile, kb0, kb1))
        it += kb1 - kb0
    return out


def owner_last(tile, G, sk_tiles, kblocks):
    S = sk_tiles * kblocks
    t0, t1 = tile * kblocks, (tile + 1) * kblocks
    return ((t0 + 1) * G + S - 1) // S - 1, (t1 * G + S - 1) // S - 1


CASES = [(1, 65536, 148), (2, 65536, 444), (1, 7500, 296), (4, 4375, 740), (3, 1000, 148), (5, 333, 444),
         (148, 32, 148), (100, 40, 296), (7, 129, 740), (1, 148, 148), (1, 149, 148), (2, 9, 148)]
random.seed(7)
CASES += [(random.randint(1, 40), random.randint(1, 5000), random.choice([148, 296, 444, 740, 132, 146 * 3]))
          for _ in range(60)]


@pytest.mark.parametrize("sk_tiles,kblocks,G", CASES)
def test_segments_cover_every_iteration_once(sk_tiles, kblocks, G):
    covered = {}
    for d in range(G):
        for (tile, kb0, kb1) in segments(d, G, sk_tiles, kblocks):
            assert 0 <= kb0 < kb1 <= kblocks and tile < sk_tiles
            for kb in {kb0, kb1 - 1}:                     # ends are enough: runs are contiguous
                assert (tile, kb) not in covered
                covered[(tile, kb)] = d
    total = sum(kb1 - kb0 for d in range(G) for (_, kb0, kb1) in segments(d, G, sk_tiles, kblocks))
    assert total == sk_tiles * kblocks


@pytest.mark.parametrize("sk_tiles,kblocks,G", CASES)
def test_owner_and_last_rank_of_each_tile(sk_tiles, kblocks, G):
    S = sk_tiles * kblocks
    if S < G:
        pytest.skip("runs can be empty: the kernel takes the linear fix-up there")
    touching = {}
    for d in range(G):
        for (tile, kb0, kb1) in segments(d, G, sk_tiles, kblocks):
            touching.setdefault(tile, []).append((d, kb0, kb1))
    for tile, parts in touching.items():
        owner, last = owner_last(tile, G, sk_tiles, kblocks)
        ranks = [d for d, _, _ in parts]
        assert ranks == list(range(owner, last + 1)), (tile, ranks, owner, last)
        assert parts[0][1] == 0 and parts[-1][2] == kblocks
        assert all(kb0 != 0 for _, kb0, _ in parts[1:])   # every non-owner publishes


@pytest.mark.parametrize("sk_tiles,kblocks,G", CASES)
def test_two_level_sum_adds_every_partial_once_and_cannot_deadlock(sk_tiles, kblocks, G):
    S = sk_tiles * kblocks
    if S < G:
        pytest.skip("linear fix-up")
    for tile in range(sk_tiles):
        owner, last = owner_last(tile, G, sk_tiles, kblocks)
        if last - owner <= SK_TREE_MIN:
            continue
        added_into = {}                                   # contributor -> who adds its partial
        for d in range(owner + 1, last + 1):              # what each non-owner does
            if (d - owner - 1) % SK_TREE_GROUP == 0:      # group leader
                for f in range(d + 1, min(d + SK_TREE_GROUP - 1, last) + 1):
                    assert f > d                          # waits point upwards only
                    assert f not in added_into
                    added_into[f] = d
        for d in range(owner + 1, last + 1, SK_TREE_GROUP):   # what the owner does
            assert (d - owner - 1) % SK_TREE_GROUP == 0       # it only reads group sums
            assert d not in added_into
            added_into[d] = owner
        assert sorted(added_into) == list(range(owner + 1, last + 1))
        # followers never wait: nobody is both a follower and a leader
        leaders = {d for d in range(owner + 1, last + 1) if (d - owner - 1) % SK_TREE_GROUP == 0}
        followers = {f for f, who in added_into.items() if who != owner}
        assert not (leaders & followers)
