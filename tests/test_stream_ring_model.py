"""Happens-before model of the tensor-stream engine's ring protocol (eigencuda_b200/csrc/stream_run.inl:
feeder, drainer; stream_engine.inl: the three streams and the per-slot events), checked on the CPU.

compute-sanitizer's racecheck is closed on the GPU pool, so the slot-reuse rules are proved here instead:
the model issues exactly the calls the feeder and the drainer issue, with CUDA's ordering semantics

  * operations of one stream run in issue order;
  * cudaStreamWaitEvent orders everything issued to the stream afterwards behind the event's LATEST record
    at the time of the call; cudaEventSynchronize does the same for the calling host thread;
  * a device operation starts after the host call that enqueued it; host threads run in program order;
    the mutex/condition-variable hand-offs between feeder and drainer order what they guard;

and then requires, for every ring buffer (pinned input, device input, device output, pinned output), that any
two accesses of which one is a write are ordered, and that they alternate write(c), read(c), write(c + ring),
read(c + ring), ...  Each wait of the protocol is also removed in turn: the checker must then find a hazard
(no wait is decorative, and the checker is not blind).
"""
import itertools

import pytest

WAITS = ("host_sync_h2d", "h2d_wait_cmp", "cmp_wait_h2d", "cmp_wait_d2h", "d2h_wait_cmp", "feeder_wait_drained",
         "drainer_wait_enqueued", "drainer_sync_d2h")


class Graph:
    def __init__(self):
        self.pred = []          # node -> set of direct predecessors
        self.acc = {}           # buffer -> list of (node, "R"/"W", chunk)

    def node(self, *after):
        self.pred.append({a for a in after if a is not None})
        return len(self.pred) - 1

    def access(self, node, buf, mode, chunk):
        self.acc.setdefault(buf, []).append((node, mode, chunk))

    def closure(self):
        """ancestors[n] as a bit mask; nodes are created in an order compatible with the edges."""
        anc = [0] * len(self.pred)
        for n, ps in enumerate(self.pred):
            m = 0
            for p in ps:
                assert p < n
                m |= anc[p] | (1 << p)
            anc[n] = m
        return anc


def build(nchunks, ring, staged_in, staged_out, drop=None):
    g = Graph()
    on = lambda w: w != drop
    last = {"feeder": None, "drainer": None, "h2d": None, "cmp": None, "d2h": None}
    pending = {"h2d": set(), "cmp": set(), "d2h": set()}      # event records the stream's next ops must follow
    rec = {}                                                   # event name -> node of its latest record

    def host(thread, *extra):
        n = g.node(last[thread], *extra)
        last[thread] = n
        return n

    def enqueue(stream, *extra):
        """A device operation issued by the feeder now."""
        issue = host("feeder")
        n = g.node(issue, last[stream], *pending[stream], *extra)
        last[stream] = n
        return n

    def record(stream, ev):
        host("feeder")
        rec[ev] = last[stream]            # completes when everything issued to the stream so far has

    def stream_wait(stream, ev):
        host("feeder")
        if ev in rec and rec[ev] is not None:
            pending[stream].add(rec[ev])

    signal_enq, signal_drained = {}, {}
    drainer_plan = []                     # the drainer's steps are interleaved below, where its inputs exist
    for c in range(nchunks):
        s = c % ring
        if staged_in:
            if c >= ring and on("host_sync_h2d"):
                host("feeder", rec[f"h2d{s}"])                          # cudaEventSynchronize(ev_h2d[s])
            n = host("feeder")                                          # stage chunk: caller -> pinned
            g.access(n, f"pin_in{s}", "W", c)
        if c >= ring and on("h2d_wait_cmp"):
            stream_wait("h2d", f"cmp{s}")
        n = enqueue("h2d")
        if staged_in:
            g.access(n, f"pin_in{s}", "R", c)
        g.access(n, f"d_in{s}", "W", c)
        record("h2d", f"h2d{s}")
        if on("cmp_wait_h2d"):
            stream_wait("cmp", f"h2d{s}")
        if c >= ring and on("cmp_wait_d2h"):
            stream_wait("cmp", f"d2h{s}")
        n = enqueue("cmp")
        g.access(n, f"d_in{s}", "R", c)
        g.access(n, f"d_out{s}", "W", c)
        record("cmp", f"cmp{s}")
        if on("d2h_wait_cmp"):
            stream_wait("d2h", f"cmp{s}")
        if staged_out and c >= ring and on("feeder_wait_drained"):
            # cv.wait(drained >= c - ring + 1): run the drainer up to that chunk first
            while len(signal_drained) < c - ring + 1:
                drain_one(g, last, rec, signal_enq, signal_drained, ring, on, host)
            host("feeder", signal_drained[c - ring])
        n = enqueue("d2h")
        g.access(n, f"d_out{s}", "R", c)
        if staged_out:
            g.access(n, f"pin_out{s}", "W", c)
        record("d2h", f"d2h{s}")
        if staged_out:
            signal_enq[c] = (host("feeder"), rec[f"d2h{s}"])            # enqueued = c + 1, and the record it announces
    if staged_out:
        while len(signal_drained) < nchunks:
            drain_one(g, last, rec, signal_enq, signal_drained, ring, on, host)
    return g


def drain_one(g, last, rec, signal_enq, signal_drained, ring, on, host):
    c = len(signal_drained)
    s = c % ring
    sig, d2h_record = signal_enq[c]
    if on("drainer_wait_enqueued"):
        host("drainer", sig)                                            # cv.wait(enqueued > c)
    if on("drainer_sync_d2h") and on("drainer_wait_enqueued"):
        # cudaEventSynchronize(ev_d2h[s]).  It only means "chunk c has arrived" because the hand-off above
        # guarantees that the record of chunk c has been issued: on an event whose record is still to come
        # the call returns at once (or waits for the slot's previous record).
        host("drainer", d2h_record)
    n = host("drainer")                                                 # deliver chunk: pinned -> caller
    g.access(n, f"pin_out{s}", "R", c)
    signal_drained[c] = host("drainer")                                 # drained = c + 1


def hazards(g):
    anc = g.closure()
    before = lambda a, b: (anc[b] >> a) & 1
    found = []
    for buf, accs in g.acc.items():
        for (n1, m1, c1), (n2, m2, c2) in itertools.combinations(accs, 2):
            if "W" in (m1, m2) and not (before(n1, n2) or before(n2, n1)):
                found.append((buf, (m1, c1), (m2, c2)))
        if not found:
            order = sorted(accs, key=lambda a: bin(anc[a[0]]).count("1"))   # a total order here: sort by depth
            want = [(m, c) for c in sorted({c for _, _, c in accs}) for m in ("W", "R")]
            if [(m, c) for _, m, c in order] != want:
                found.append((buf, "order", [(m, c) for _, m, c in order]))
    return found


CONFIGS = [(nchunks, ring, si, so) for nchunks in (1, 2, 3, 5, 9, 13) for ring in (2, 3, 4, 5)
           for si in (False, True) for so in (False, True)]


@pytest.mark.parametrize("nchunks,ring,staged_in,staged_out", CONFIGS)
def test_ring_protocol_has_no_hazard(nchunks, ring, staged_in, staged_out):
    assert hazards(build(nchunks, ring, staged_in, staged_out)) == []


@pytest.mark.parametrize("drop", WAITS)
def test_every_wait_is_needed(drop):
    """Remove one wait: some configuration must now show an unordered pair of accesses."""
    seen = False
    for nchunks, ring, si, so in CONFIGS:
        if nchunks <= ring:
            continue
        if drop in ("host_sync_h2d",) and not si:
            continue
        if drop in ("feeder_wait_drained", "drainer_wait_enqueued", "drainer_sync_d2h") and not so:
            continue
        try:
            bad = hazards(build(nchunks, ring, si, so, drop=drop))
        except KeyError:      # the feeder went on without the drainer: its hand-off never happened
            bad = ["feeder ran ahead of the drainer"]
        seen = seen or bool(bad)
    assert seen, f"dropping {drop} went unnoticed"
