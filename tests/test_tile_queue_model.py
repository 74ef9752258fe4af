"""Model of the whole-tile part of the work schedule (eigencuda_b200/csrc/dgemm_kernels.cuh: the producer's
claim loop and tile queue, the consumers' pops; csrc/dgemm_launch.inl: the host's bookkeeping of the claim
counter), run on the CPU under random interleavings of the CTAs.  What is checked:

  * every whole tile is loaded by exactly one CTA, and its consumers compute exactly the tiles its producer
    loaded, in the same order (the fixed first tile is not queued, everything after it is, -1 ends the stream);
  * the producer never runs more than TQ queue entries ahead of the consumers;
  * the number of claims a launch makes on the never-reset global counter is what the host adds to
    `dp_claims` BEFORE the launch (dp_tiles, or dp_tiles + grid when every tile is claimed) -- if that were
    off by one, the next launch on the pipeline would skip or repeat a tile.
"""
import random

import pytest

TQ = 4   # dgemm_kernels.cuh


class Launch:
    """Host side of one launch (launch_tma_dmma): what the kernel is told, what the counter is advanced by."""

    def __init__(self, counter_before, total_tiles, sk_tiles, grid, kblocks, dynamic, all_dynamic):
        self.total, self.sk, self.grid, self.kblocks = total_tiles, sk_tiles, grid, kblocks
        dp_tiles = total_tiles - sk_tiles
        self.dp_counter, self.all_dynamic, self.dp_base, self.predicted = False, False, 0, 0
        if dynamic and all_dynamic and dp_tiles > 0:
            self.dp_counter, self.all_dynamic, self.dp_base = True, True, counter_before
            self.predicted = dp_tiles + grid
        elif dynamic and dp_tiles > grid and kblocks >= 8:
            self.dp_counter, self.dp_base = True, counter_before
            self.predicted = dp_tiles
        self.queued = self.all_dynamic or dp_tiles > grid


def producer(L, cta, counter, loads, queue):
    """Generator: the elected producer thread of one CTA; yields at every point where it may be descheduled
    or has to wait (claims, queue pushes, tile loads)."""
    dyn_first = L.sk + (0 if L.all_dynamic else L.grid)

    def claim():
        v = counter[0]
        counter[0] += 1
        return dyn_first + (v - L.dp_base)

    first_static = not L.all_dynamic
    if first_static:
        cur = L.sk + cta
    else:
        cur = claim()
        yield
    first = first_static
    while True:
        valid = cur < L.total
        if L.queued and not (first and valid):
            while len(queue) >= TQ:          # mbar_wait(tq_empty)
                yield
            queue.append(cur if valid else -1)
            yield
        if not valid:
            break
        nxt = L.total                        # "invalid" until claimed
        claim_kb = (L.kblocks - 8 if L.kblocks > 8 else 0) if L.queued else -1
        for kb in range(L.kblocks):          # load_segment
            if kb == claim_kb:
                nxt = claim() if L.dp_counter else cur + L.grid
                yield
        loads.append(cur)
        yield
        if not L.queued:
            break
        cur = nxt
        first = False


def consumer(L, cta, computed, queue):
    qn = 0 if L.all_dynamic else -2
    while True:
        if qn == -2:
            qn = 0 if L.queued else -1
            if L.sk + cta >= L.total:
                continue
            computed.append(L.sk + cta)
            yield
        elif qn >= 0:
            while not queue:                 # mbar_wait(tq_full)
                yield
            t = queue.pop(0)
            if t < 0:
                qn = -1
                continue
            qn = (qn + 1) % (2 * TQ)
            computed.append(t)
            yield
        else:
            break


def run_launch(L, counter, rng):
    loads = [[] for _ in range(L.grid)]
    computed = [[] for _ in range(L.grid)]
    queues = [[] for _ in range(L.grid)]
    threads = []
    for c in range(L.grid):
        threads.append(producer(L, c, counter, loads[c], queues[c]))
        threads.append(consumer(L, c, computed[c], queues[c]))
    steps = 0
    while threads:
        i = rng.randrange(len(threads))
        try:
            next(threads[i])
        except StopIteration:
            threads.pop(i)
        steps += 1
        assert steps < 5_000_000, "the model deadlocked"
        assert all(len(q) <= TQ for q in queues)
    return loads, computed


CASES = [  # total tiles, stream-K tiles, grid, k-blocks
    (1024, 0, 148, 32), (300, 4, 148, 32), (149, 0, 148, 64), (148, 0, 148, 32), (100, 0, 148, 32),
    (2 * 148, 0, 148, 32), (2 * 148 + 1, 1 + 148, 148, 16), (500, 0, 148, 4), (500, 0, 148, 8), (37, 5, 16, 9),
    (64, 0, 8, 1), (1000, 0, 296, 12),
]


@pytest.mark.parametrize("total,sk,grid,kblocks", CASES)
@pytest.mark.parametrize("mode", ["static", "dynamic", "all_dynamic"])
def test_tile_queue_and_claim_counter(total, sk, grid, kblocks, mode):
    rng = random.Random(f"{total}-{sk}-{grid}-{kblocks}-{mode}")   # a string seed is stable across processes
    counter = [rng.randrange(0, 1000)]       # the counter is never reset: it starts wherever the last launch left it
    for _ in range(3):                       # back-to-back launches on one pipeline share the counter
        before = counter[0]
        L = Launch(before, total, sk, grid, kblocks, mode != "static", mode == "all_dynamic")
        loads, computed = run_launch(L, counter, rng)
        assert computed == loads                                         # consumers follow their producer
        everything = sorted(t for per_cta in loads for t in per_cta)
        assert everything == list(range(sk, total))                      # each whole tile exactly once
        assert counter[0] - before == L.predicted                        # the host's bookkeeping holds
        if mode == "static" or not L.dp_counter:
            for c in range(grid):                                        # round-robin: sk + c, sk + c + grid, ...
                assert loads[c] == list(range(sk + c, total, grid))
