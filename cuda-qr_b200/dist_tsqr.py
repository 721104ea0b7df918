"""Row-partitioned TSQR across the GPUs of one box (BASELINE config 3).

Each rank owns a contiguous block of rows (column-major local slab, lda = m_local), runs a local
TSQR through the C ABI (``cqr_tsqr_r`` / ``cqr_tsqr_factor``) and the n x n R factors are combined
in a binary reduction tree whose stacked-R blocks move with point-to-point messages
(``torch.distributed`` send/recv = ncclSend/ncclRecv over NVLink): at level s the rank with bit s
set sends its R to rank - 2^s, which stacks [R_mine; R_recv] and re-factors it (``cqr_stack_qr``).
log2(P) hops of 16 KiB each; no collective on the data path.  The reference has no multi-GPU code
(qr.cu:737), so parity is judged against the single-device / oracle result on the same input.

The tree schedule is plain rank arithmetic and is kept free of CUDA so the gloo CPU tests can run it
with an injected combine step.
"""
from __future__ import annotations

from typing import Callable, List, Tuple


def rtree_steps(rank: int, world: int) -> List[Tuple[str, int, int]]:
    """Actions of `rank` in the reduction tree, in order: ("recv", peer, level) merges peer's R into
    ours; ("send", peer, level) hands ours over and ends this rank's part.  Works for any world >= 1."""
    steps = []
    s, level = 1, 0
    while s < world:
        if rank % (2 * s) == s:
            steps.append(("send", rank - s, level))
            break
        if rank % (2 * s) == 0 and rank + s < world:
            steps.append(("recv", rank + s, level))
        s *= 2
        level += 1
    return steps


def tree_depth(world: int) -> int:
    d, s = 0, 1
    while s < world:
        s *= 2
        d += 1
    return d


def reduce_r(rank: int, world: int, r_local, combine: Callable, send: Callable, recv: Callable):
    """Generic R-tree: `combine(r_mine, r_peer, level)` returns the merged factor.  Returns the final R on
    rank 0 and None elsewhere."""
    r = r_local
    for kind, peer, level in rtree_steps(rank, world):
        if kind == "send":
            send(r, peer)
            return None
        r = combine(r, recv(peer), level)
    return r if rank == 0 else None


class DistTSQR:
    """TSQR of a row-partitioned tall-skinny matrix; one instance per rank (one process per GPU)."""

    def __init__(self, pkg, ctx, n: int, rank: int, world: int, device, comm=None):
        import torch
        self.torch, self.pkg, self.ctx = torch, pkg, ctx
        self.n, self.rank, self.world, self.device = n, rank, world, device
        if comm is None and world > 1:
            from .comm import Comm
            comm = Comm()
        self.comm = comm                     # transport only (NCCL; gloo stages through the host, see comm.py)
        depth = tree_depth(world)
        # per tree level: the stacked [R_mine; R_recv] tile (kept: it holds the level's reflectors) and its tau
        self.stack = [pkg.colmajor(2 * n, n, device=device) for _ in range(depth)]
        self.tau = [torch.zeros(64, device=device) for _ in range(depth)]
        self.rbuf = pkg.colmajor(n, n, device=device)
        self.R = pkg.colmajor(n, n, device=device)
        self.steps = rtree_steps(rank, world)
        self.peer = False                    # True after enable_peer(): the R-only tree runs over peer memory inside the library

    def enable_peer(self):
        """Switch the R-only path to the library's peer-memory R tree (cqr_dist_*, rtree_peer.cu): every rank exports the
        cudaIpc handle of its exchange slab, the 64-byte handles are gathered through torch.distributed (any backend), and
        from then on factor(keep_q=False) is ONE library call per rank -- local TSQR plus a tree kernel whose hand-overs are
        NVLink stores into the receiver's slab -- with no NCCL message and no host synchronisation between calls."""
        if self.world == 1:
            return self
        import torch.distributed as dist
        mine = self.ctx.dist_export()
        handles = [None] * self.world
        dist.all_gather_object(handles, mine)
        self.ctx.dist_attach(self.rank, self.world, handles)
        self.peer = True
        return self

    @staticmethod
    def _wire(t):
        """colmajor() tensors are transposed views of contiguous (n, ld) storage: send that storage."""
        base = t.t()
        assert base.is_contiguous(), "R-tree messages need ld == rows"
        return base

    def factor(self, A_local, keep_q: bool = False):
        """Local TSQR + R-tree.  The final R is valid on rank 0 (self.R)."""
        n = self.n
        if self.peer and not keep_q:
            self.ctx.tsqr_dist_r(A_local, self.R)
            return self.R
        if keep_q:
            self.ctx.tsqr_factor(A_local, self.R)
        else:
            self.ctx.tsqr_r(A_local, self.R)
        for kind, peer, level in self.steps:
            if kind == "send":
                self.comm.send(self._wire(self.R), peer)
                break
            st = self.stack[level]
            st[:n].copy_(self.R)
            self.comm.recv(self._wire(self.rbuf), peer)
            st[n:].copy_(self.rbuf)
            self.ctx.stack_qr(st, n, self.tau[level], self.R)
        return self.R

    def broadcast_r(self):
        if self.world > 1:
            self.comm.broadcast(self._wire(self.R), 0)
        return self.R

    def form_q(self, Q_local):
        """Explicit thin Q rows of this rank (after factor(keep_q=True)): walk the tree root -> leaves, each
        combine node expanding its seed X into [X_mine; X_peer] = Q_node [X; 0] and sending X_peer down."""
        n = self.n
        X = None
        # ranks that sent their R receive their seed from the parent first
        mine = list(self.steps)
        if mine and mine[-1][0] == "send":
            X = self.pkg.colmajor(n, n, device=self.device)
            self.comm.recv(self._wire(X), mine[-1][1])
            mine = mine[:-1]
        for kind, peer, level in reversed(mine):
            Qs = self.pkg.colmajor(2 * n, n, device=self.device)
            self.ctx.stack_form_q(self.stack[level], n, self.tau[level], Qs, X)
            Xp = self.pkg.colmajor(n, n, device=self.device)
            Xp.copy_(Qs[n:])
            self.comm.send(self._wire(Xp), peer)
            X = self.pkg.colmajor(n, n, device=self.device)
            X.copy_(Qs[:n])
        self.ctx.tsqr_form_q(Q_local, X)
        return Q_local
