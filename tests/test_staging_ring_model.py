"""Happens-before model of the pinned staging ring behind the reference's blocking transfers
(eigencuda_b200/csrc/host_util.inl: StagingRing; ec_api.cu: ec_matrix_upload, ec_matrix_download), on the CPU.

copy_to_gpu must leave its pageable source reusable on return (reference src/cudamatrix.cc:47-51) while the DMA
itself runs later from a pinned slot; the conversion back DMAs into two slots alternately and copies out behind
the DMA.  Slots rotate over SLOTS buffers: a slot may only be overwritten by the host (or by the next DMA) once
the DMA that last used it has completed -- `acquire` synchronises on the slot's event when it is still busy."""
import pytest

from tests.hb_model import Cuda

SLOTS = 4   # StagingRing::SLOTS


class Ring:
    def __init__(self, g, sync_when_busy=True):
        self.g, self.next, self.busy, self.sync = g, 0, [False] * SLOTS, sync_when_busy

    def acquire(self):
        s = self.next
        self.next = (self.next + 1) % SLOTS
        if self.busy[s]:
            if self.sync:
                self.g.event_sync(f"slot{s}")
            self.busy[s] = False
        return s

    def release_after(self, s, stream):
        self.g.record(stream, f"slot{s}")
        self.busy[s] = True


def upload(g, ring, tag, pieces):
    """ec_matrix_upload of a pageable matrix: per piece acquire -> host copy into the slot -> DMA -> release_after."""
    for _ in range(pieces):
        s = ring.acquire()
        g.access(g.host(), f"pin{s}", "W", tag[0])            # compat_copy: caller -> pinned
        g.access(g.enqueue("stream"), f"pin{s}", "R", tag[0])  # cudaMemcpyAsync H2D from the slot
        ring.release_after(s, "stream")
        tag[0] += 1


def download(g, ring, tag, pieces, sync_before_copy_out=True):
    """ec_matrix_download into pageable memory: DMA piece i into a slot, copy piece i-1 out meanwhile."""
    prev = None
    for i in range(pieces + 1):
        cur = None
        if i < pieces:
            s = ring.acquire()
            g.access(g.enqueue("stream"), f"pin{s}", "W", tag[0])   # cudaMemcpyAsync D2H into the slot
            ring.release_after(s, "stream")
            cur = (s, tag[0])
            tag[0] += 1
        if prev is not None:
            s, t = prev
            if sync_before_copy_out:
                g.event_sync(f"slot{s}")
            ring.busy[s] = False
            g.access(g.host(), f"pin{s}", "R", t)                   # compat_copy: pinned -> caller
        prev = cur


@pytest.mark.parametrize("pattern", ["uploads", "downloads", "readme_loop", "large_matrices"])
def test_staging_slots_are_never_reused_under_a_running_dma(pattern):
    g = Cuda(streams=("stream",))
    ring, tag = Ring(g), [0]
    if pattern == "uploads":
        for _ in range(11):
            upload(g, ring, tag, 1)
    elif pattern == "downloads":
        for _ in range(5):
            download(g, ring, tag, 2)
    elif pattern == "readme_loop":                 # copy_to_gpu -> gemm -> convert, README.md:62-66
        for _ in range(7):
            upload(g, ring, tag, 1)
            g.enqueue("stream")                    # the product
            download(g, ring, tag, 2)
    else:                                          # matrices above 8 MiB go in 4 MiB pieces over the rotating slots
        upload(g, ring, tag, 9)
        download(g, ring, tag, 9)
        upload(g, ring, tag, 3)
    assert g.hazards() == []


def test_the_model_notices_a_missing_synchronisation():
    g = Cuda(streams=("stream",))
    ring, tag = Ring(g, sync_when_busy=False), [0]
    for _ in range(SLOTS + 1):
        upload(g, ring, tag, 1)
    assert g.hazards()                             # the fifth upload overwrites slot 0 under the first DMA
    g = Cuda(streams=("stream",))
    ring, tag = Ring(g), [0]
    download(g, ring, tag, 2, sync_before_copy_out=False)
    assert g.hazards()                             # copied out before the DMA has landed
