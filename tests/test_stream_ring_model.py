"""Happens-before model of the tensor-stream engine's ring protocol (eigencuda_b200/csrc/stream_run.inl:
feeder, drainer; stream_engine.inl: the three streams and the per-slot events), checked on the CPU.

compute-sanitizer's racecheck is closed on the GPU pool, so the slot-reuse rules are proved here instead:
the model issues exactly the calls the feeder and the drainer issue, with CUDA's ordering semantics
(tests/hb_model.py) plus the mutex / condition-variable hand-offs between the two host threads, and then requires,
for every ring buffer (pinned input, device input, device output, pinned output), that any two accesses of which
one is a write are ordered, and that chunk c's data is written before it is read and read before chunk c + ring
overwrites it.  Each wait of the protocol is also removed in turn: the checker must then find a hazard (no wait is
decorative, and the checker is not blind).
"""
import pytest

from tests.hb_model import Cuda

WAITS = ("host_sync_h2d", "h2d_wait_cmp", "cmp_wait_h2d", "cmp_wait_d2h", "d2h_wait_cmp", "feeder_wait_drained",
         "drainer_wait_enqueued", "drainer_sync_d2h")
F, D = "feeder", "drainer"


def build(nchunks, ring, staged_in, staged_out, drop=None):
    on = lambda w: w != drop
    g = Cuda(streams=("h2d", "cmp", "d2h"), host_threads=(F, D))
    announced, drained = {}, {}       # chunk -> (feeder's "enqueued = c + 1" node, its D2H record) / drainer's "drained = c + 1" node

    def drain_next():
        """The drainer's next chunk (its steps are interleaved where their inputs exist)."""
        c = len(drained)
        s = c % ring
        sig, d2h_done = announced[c]
        if on("drainer_wait_enqueued"):
            g.host(D, sig)                                              # cv.wait(enqueued > c)
        if on("drainer_sync_d2h") and on("drainer_wait_enqueued"):
            # cudaEventSynchronize(ev_d2h[s]).  It only means "chunk c has arrived" because the hand-off above
            # guarantees that the record of chunk c has been issued: on an event whose record is still to come
            # the call returns at once (or waits for the slot's previous record).
            g.host(D, d2h_done)
        g.access(g.host(D), f"pin_out{s}", "R", c)                      # deliver chunk: pinned -> caller
        drained[c] = g.host(D)                                          # drained = c + 1

    for c in range(nchunks):
        s = c % ring
        if staged_in:
            if c >= ring and on("host_sync_h2d"):
                g.event_sync(f"h2d{s}", F)                              # pin_in[s] consumed
            g.access(g.host(F), f"pin_in{s}", "W", c)                   # stage chunk: caller -> pinned
        if c >= ring and on("h2d_wait_cmp"):
            g.stream_wait("h2d", f"cmp{s}", F)                          # d_in[s] consumed
        n = g.enqueue("h2d", F)
        if staged_in:
            g.access(n, f"pin_in{s}", "R", c)
        g.access(n, f"d_in{s}", "W", c)
        g.record("h2d", f"h2d{s}", F)
        if on("cmp_wait_h2d"):
            g.stream_wait("cmp", f"h2d{s}", F)
        if c >= ring and on("cmp_wait_d2h"):
            g.stream_wait("cmp", f"d2h{s}", F)                          # d_out[s] drained
        n = g.enqueue("cmp", F)
        g.access(n, f"d_in{s}", "R", c)
        g.access(n, f"d_out{s}", "W", c)
        g.record("cmp", f"cmp{s}", F)
        if on("d2h_wait_cmp"):
            g.stream_wait("d2h", f"cmp{s}", F)
        if staged_out and c >= ring and on("feeder_wait_drained"):
            while len(drained) < c - ring + 1:                          # cv.wait(drained >= c - ring + 1)
                drain_next()
            g.host(F, drained[c - ring])
        n = g.enqueue("d2h", F)
        g.access(n, f"d_out{s}", "R", c)
        if staged_out:
            g.access(n, f"pin_out{s}", "W", c)
        g.record("d2h", f"d2h{s}", F)
        if staged_out:
            announced[c] = (g.host(F), g.rec[f"d2h{s}"])                # enqueued = c + 1
    while staged_out and len(drained) < nchunks:
        drain_next()
    return g


CONFIGS = [(nchunks, ring, si, so) for nchunks in (1, 2, 3, 5, 9, 13) for ring in (2, 3, 4, 5)
           for si in (False, True) for so in (False, True)]


@pytest.mark.parametrize("nchunks,ring,staged_in,staged_out", CONFIGS)
def test_ring_protocol_has_no_hazard(nchunks, ring, staged_in, staged_out):
    assert build(nchunks, ring, staged_in, staged_out).hazards() == []


@pytest.mark.parametrize("drop", WAITS)
def test_every_wait_is_needed(drop):
    """Remove one wait: some configuration must now show an unordered pair of accesses."""
    seen = False
    for nchunks, ring, si, so in CONFIGS:
        if nchunks <= ring or (drop == "host_sync_h2d" and not si):
            continue
        if drop in ("feeder_wait_drained", "drainer_wait_enqueued", "drainer_sync_d2h") and not so:
            continue
        seen = seen or bool(build(nchunks, ring, si, so, drop=drop).hazards())
    assert seen, f"dropping {drop} went unnoticed"
