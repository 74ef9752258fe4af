"""A small happens-before calculator for protocols written against CUDA streams and events (test
infrastructure for tests/test_stream_ring_model.py and tests/test_summa_model.py).

Semantics modelled: operations of one stream run in issue order; cudaStreamWaitEvent orders whatever is issued to
the stream afterwards behind the event's LATEST record at the time of the call (a wait on a never-recorded event
is a no-op); a device operation starts after the host call that enqueued it; a host thread runs in program order;
cudaEventSynchronize orders the calling host thread behind the record.  Accesses to named buffers are logged so
that a protocol can be checked for unordered conflicting pairs."""
import itertools


class Cuda:
    def __init__(self, streams, host_threads=("host",)):
        self.pred = []                                   # node -> set of direct predecessors (always older nodes)
        self.acc = {}                                    # buffer -> [(node, "R"/"W", tag)]
        self.last = {name: None for name in tuple(streams) + tuple(host_threads)}
        self.pending = {s: set() for s in streams}       # records the stream's future operations must follow
        self.rec = {}                                    # event -> node its latest record stands for

    def _node(self, *after):
        self.pred.append({a for a in after if a is not None})
        return len(self.pred) - 1

    def host(self, thread="host", *after):
        """A host-side step of `thread` (optionally ordered behind other nodes: a cv hand-off, an event sync)."""
        n = self._node(self.last[thread], *after)
        self.last[thread] = n
        return n

    def enqueue(self, stream, thread="host"):
        issue = self.host(thread)
        n = self._node(issue, self.last[stream], *self.pending[stream])
        self.last[stream] = n
        return n

    def record(self, stream, event, thread="host"):
        self.host(thread)
        self.rec[event] = self.last[stream]              # complete when everything issued to the stream so far is

    def stream_wait(self, stream, event, thread="host"):
        self.host(thread)
        if self.rec.get(event) is not None:
            self.pending[stream].add(self.rec[event])

    def event_sync(self, event, thread="host"):
        return self.host(thread, self.rec.get(event))

    def access(self, node, buf, mode, tag):
        self.acc.setdefault(buf, []).append((node, mode, tag))

    def ancestors(self):
        anc = [0] * len(self.pred)
        for n, ps in enumerate(self.pred):
            m = 0
            for p in ps:
                assert p < n
                m |= anc[p] | (1 << p)
            anc[n] = m
        return anc

    def hazards(self, expect_alternation=True):
        """Unordered conflicting pairs; and, when every pair is ordered, whether each buffer sees
        W(tag0) R(tag0)... W(tag1) R(tag1)... in tag order (all readers of a tag between its write and the next)."""
        anc = self.ancestors()
        before = lambda a, b: (anc[b] >> a) & 1
        found = []
        for buf, accs in self.acc.items():
            for (n1, m1, t1), (n2, m2, t2) in itertools.combinations(accs, 2):
                if "W" in (m1, m2) and n1 != n2 and not (before(n1, n2) or before(n2, n1)):
                    found.append((buf, (m1, t1), (m2, t2)))
            if expect_alternation and not found:
                for (n1, m1, t1), (n2, m2, t2) in itertools.permutations(accs, 2):
                    if n1 == n2:                  # one operation that reads and rewrites the buffer
                        continue
                    if m1 == "W" and m2 == "R" and t1 == t2 and not before(n1, n2):
                        found.append((buf, "read before its write", t1))
                    if m1 == "R" and m2 == "W" and t1 < t2 and not before(n1, n2):
                        found.append((buf, "overwritten before read", t1, t2))
        return found
