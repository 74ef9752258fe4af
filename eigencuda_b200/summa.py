"""2D block-partitioned DGEMM across the GPUs of one box (BASELINE.json configs[4]: n = 65536 on
8 B200).  One process per GPU; the only exchange on the path is the panel
broadcast of SUMMA, done with NCCL over NVLink/NVSwitch (torch.distributed) on a side stream and
double-buffered against the local product, which is this package's DGEMM kernel with beta = 1
accumulating into the resident C block.  The reference has nothing like this (single device,
dimensions narrowed to int: reference src/cudapipeline.cc:29-31, Appendix A #4).

Layout (all column-major, like everything else here).  P x Q process grid, rank r at
(pr, pc) = (r // Q, r % Q).  For C(m x n) = A(m x k) * B(k x n):

    A is cut into P row blocks x Q column blocks, rank (pr, pc) holds A[pr, pc]  (m/P x k/Q)
    B is cut into P row blocks x Q column blocks, rank (pr, pc) holds B[pr, pc]  (k/P x n/Q)
    C likewise,                                   rank (pr, pc) holds C[pr, pc]  (m/P x n/Q)

For every k-panel [k0, k0 + kb): the grid column that owns those columns of A broadcasts its
A[pr, :][:, panel] (m/P x kb, contiguous in column-major storage) along each grid row; the grid
row that owns those rows of B broadcasts B[:, pc][panel, :] (kb x n/Q, packed) along each grid
column; every rank then does C[pr, pc] += Apanel * Bpanel.

Why not a kernel that TMA-loads the panels straight from the peers: peer addresses bypass the
local L2 (B300_MICROARCH.md, NVLink section), so every CTA's re-read of a panel tile would cross
NVLink -- about 1.2 TB/s for this kernel against 0.77 TB/s available -- whereas a broadcast into
local HBM costs each byte one crossing (17 GB per GPU at n = 65536, ~25 ms against ~1.9 s of DMMA).
"""
import math


def grid_shape(world):
    """P x Q with P <= Q, as square as possible: 1x1, 1x2, 2x2, 2x4."""
    p = int(math.isqrt(world))
    while world % p:
        p -= 1
    return p, world // p


def block_range(total, parts, idx):
    lo = (total * idx) // parts
    hi = (total * (idx + 1)) // parts
    return lo, hi


class Layout:
    """Block ownership for one rank."""

    def __init__(self, m, n, k, world, rank, grid=None):
        self.m, self.n, self.k = m, n, k
        self.P, self.Q = grid or grid_shape(world)
        if self.P * self.Q != world:
            raise ValueError("grid does not match world size")
        self.rank, self.pr, self.pc = rank, rank // self.Q, rank % self.Q
        self.m0, self.m1 = block_range(m, self.P, self.pr)        # rows of A / C held here
        self.n0, self.n1 = block_range(n, self.Q, self.pc)        # columns of B / C held here
        self.ka0, self.ka1 = block_range(k, self.Q, self.pc)      # columns of A held here
        self.kb0, self.kb1 = block_range(k, self.P, self.pr)      # rows of B held here

    @property
    def a_shape(self):
        return self.m1 - self.m0, self.ka1 - self.ka0

    @property
    def b_shape(self):
        return self.kb1 - self.kb0, self.n1 - self.n0

    @property
    def c_shape(self):
        return self.m1 - self.m0, self.n1 - self.n0

    def panels(self, kb):
        """k-panels [k0, k1) that never straddle an ownership boundary of A's columns or B's rows."""
        cuts = {0, self.k}
        for q in range(self.Q + 1):
            cuts.add(block_range(self.k, self.Q, min(q, self.Q - 1))[0] if q < self.Q else self.k)
        for p in range(self.P + 1):
            cuts.add(block_range(self.k, self.P, min(p, self.P - 1))[0] if p < self.P else self.k)
        cuts = sorted(cuts)
        out = []
        for lo, hi in zip(cuts, cuts[1:]):
            s = lo
            while s < hi:
                e = min(s + kb, hi)
                out.append((s, e))
                s = e
        return out

    def a_owner_col(self, k0):
        for q in range(self.Q):
            lo, hi = block_range(self.k, self.Q, q)
            if lo <= k0 < hi:
                return q, lo
        raise ValueError(k0)

    def b_owner_row(self, k0):
        for p in range(self.P):
            lo, hi = block_range(self.k, self.P, p)
            if lo <= k0 < hi:
                return p, lo
        raise ValueError(k0)


def make_groups(layout):
    """Row and column sub-communicators; every rank creates every group, in the same order."""
    import torch.distributed as dist
    row_groups, col_groups = [], []
    for pr in range(layout.P):
        row_groups.append(dist.new_group([pr * layout.Q + q for q in range(layout.Q)]))
    for pc in range(layout.Q):
        col_groups.append(dist.new_group([p * layout.Q + pc for p in range(layout.P)]))
    return row_groups[layout.pr], col_groups[layout.pc]


def summa(layout, A_loc, B_loc, C_loc, local_gemm, kb=1024, row_group=None, col_group=None,
          comm_stream=None):
    """C_loc = A * B for this rank's block.

    A_loc, B_loc, C_loc: torch tensors holding the local blocks in COLUMN-MAJOR order as flat
    buffers viewed (cols, rows) row-major, i.e. X_loc[j, i] == X(i, j); shapes a_shape[::-1] etc.
    local_gemm(m, n, k, A_ptr_tensor, lda, B_ptr_tensor, ldb, beta, C_tensor, ldc): the local
    product C = A*B + beta*C on column-major operands (the GPU kernel in production; the CPU
    oracle in the gloo test).  Panels are double-buffered: the broadcasts of panel t+1 run on
    `comm_stream` while panel t is multiplied.
    """
    import torch
    import torch.distributed as dist
    dev = C_loc.device
    on_gpu = dev.type == "cuda"
    mloc, nloc = layout.c_shape
    panels = layout.panels(kb)
    kbmax = max(e - s for s, e in panels)
    bufA = [torch.empty((kbmax, mloc), dtype=torch.float64, device=dev) for _ in range(2)]
    bufB = [torch.empty((nloc, kbmax), dtype=torch.float64, device=dev) for _ in range(2)]
    if on_gpu:
        main = torch.cuda.current_stream()
        comm_stream = comm_stream or torch.cuda.Stream()
        ready = [torch.cuda.Event() for _ in range(2)]     # panel landed in buffer b
        freed = [torch.cuda.Event() for _ in range(2)]     # buffer b consumed by the product

    def fetch(t, b):
        k0, k1 = panels[t]
        w = k1 - k0
        qa, a_lo = layout.a_owner_col(k0)
        pb, b_lo = layout.b_owner_row(k0)
        pa_view = bufA[b][:w]                       # (w, mloc): w columns of the A panel
        pb_view = bufB[b][:, :w]                    # (nloc, w): the B panel, ld = kbmax
        if layout.pc == qa:
            pa_view.copy_(A_loc[k0 - a_lo:k1 - a_lo])            # contiguous column range
        if layout.pr == pb:
            pb_view.copy_(B_loc[:, k0 - b_lo:k1 - b_lo])         # strided row range, packed
        if layout.Q > 1:
            dist.broadcast(pa_view, src=layout.pr * layout.Q + qa, group=row_group)
        if layout.P > 1:
            # broadcast needs a dense tensor: stage the strided view through a packed one
            packed = pb_view.contiguous()
            dist.broadcast(packed, src=pb * layout.Q + layout.pc, group=col_group)
            if packed.data_ptr() != pb_view.data_ptr():
                pb_view.copy_(packed)

    def issue_fetch(t, b):
        if on_gpu:
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(freed[b]) if t >= 2 else None
                fetch(t, b)
                ready[b].record(comm_stream)
        else:
            fetch(t, b)

    if on_gpu:
        comm_stream.wait_stream(main)
    issue_fetch(0, 0)
    for t, (k0, k1) in enumerate(panels):
        b = t & 1
        if t + 1 < len(panels):
            issue_fetch(t + 1, b ^ 1)
        if on_gpu:
            main.wait_event(ready[b])
        w = k1 - k0
        local_gemm(mloc, nloc, w, bufA[b], mloc, bufB[b], kbmax, 0.0 if t == 0 else 1.0, C_loc, mloc)
        if on_gpu:
            freed[b].record(main)
    if on_gpu:
        main.wait_stream(comm_stream)
    return C_loc
