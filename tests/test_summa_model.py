"""Happens-before model of the 2D-partitioned DGEMM's panel double buffering (eigencuda_b200/csrc/dgemm_2d.inl:
dgemm_2d_impl -- `fetch`, the panel loop, the ready / freed events), per device, on the CPU.  The panel traffic of
panel t+1 (copy-engine pack + NCCL broadcasts on the side stream) runs underneath the product of panel t (compute
stream); the two panel buffers of a device must never be overwritten while a product still reads them, and a
product must never start before its panels have landed.  Each wait is also removed in turn."""
import pytest

from tests.hb_model import Cuda

WAITS = ("comm_wait_freed", "compute_wait_ready", "comm_wait_t0")


def build(npanels, drop=None):
    on = lambda w: w != drop
    g = Cuda(streams=("compute", "comm"))
    # what is already queued on the compute stream when ec_dgemm_2d is called: the fill / upload of the blocks
    n = g.enqueue("compute")
    g.access(n, "A_block", "W", 0)
    g.access(n, "B_block", "W", 0)
    g.record("compute", "t0")
    if on("comm_wait_t0"):
        g.stream_wait("comm", "t0")          # "the side stream starts behind whatever is queued on the compute stream"

    def fetch(t, b):
        if t >= 2 and on("comm_wait_freed"):
            g.stream_wait("comm", f"freed{b}")
        n = g.enqueue("comm")                # cudaMemcpy2DAsync: pack the owner's rows of B into the panel buffer
        g.access(n, "B_block", "R", 0)
        g.access(n, f"bufB{b}", "W", 2 * t)
        n = g.enqueue("comm")                # grouped ncclBroadcast of the A panel (from the owner's block) and the B panel
        g.access(n, "A_block", "R", 0)
        g.access(n, f"bufA{b}", "W", 2 * t + 1)
        g.access(n, f"bufB{b}", "R", 2 * t)  # the root sends what the pack wrote ...
        g.access(n, f"bufB{b}", "W", 2 * t + 1)   # ... the others receive into it
        g.record("comm", f"ready{b}")

    fetch(0, 0)
    for t in range(npanels):
        b = t & 1
        if t + 1 < npanels:
            fetch(t + 1, b ^ 1)
        if on("compute_wait_ready"):
            g.stream_wait("compute", f"ready{b}")
        n = g.enqueue("compute")             # C += Apanel * Bpanel
        g.access(n, f"bufA{b}", "R", 2 * t + 1)
        g.access(n, f"bufB{b}", "R", 2 * t + 1)
        g.record("compute", f"freed{b}")
    return g


@pytest.mark.parametrize("npanels", [1, 2, 3, 4, 7, 16])
def test_panel_double_buffering_has_no_hazard(npanels):
    assert build(npanels).hazards() == []


@pytest.mark.parametrize("drop", WAITS)
def test_every_wait_is_needed(drop):
    assert any(build(n, drop=drop).hazards() for n in (1, 2, 3, 4, 7)), f"dropping {drop} went unnoticed"
