"""world_size > 1 host logic on CPU (gloo): the R-tree schedule of cuda-qr_b200/dist_tsqr.py with an injected
combine step (numpy QR stands in for the GPU's cqr_stack_qr; the product path itself is covered by -m gpu)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, m, n, out):
    import sys
    sys.path.insert(0, ROOT)
    dt = importlib.import_module("cuda-qr_b200.dist_tsqr")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    A = rng.standard_normal((m, n))
    lo, hi = rank * m // world, (rank + 1) * m // world
    r_local = np.linalg.qr(A[lo:hi], mode="r")

    def send(r, peer):
        dist.send(torch.from_numpy(np.ascontiguousarray(r)), peer)

    def recv(peer):
        t = torch.empty((n, n), dtype=torch.float64)
        dist.recv(t, peer)
        return t.numpy()

    def combine(a, b, level):
        return np.linalg.qr(np.vstack([a, b]), mode="r")

    r = dt.reduce_r(rank, world, r_local, combine, send, recv)
    if rank == 0:
        ref = np.linalg.qr(A, mode="r")
        s1, s2 = np.sign(np.diag(r)), np.sign(np.diag(ref))
        out.put(float(np.linalg.norm(r * s1[:, None] - ref * s2[:, None]) / np.linalg.norm(ref)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_rtree_reduces_to_the_global_r(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 4000, 16, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) < 1e-12


def test_rtree_schedule_properties():
    dt = importlib.import_module("cuda-qr_b200.dist_tsqr")
    for world in range(1, 17):
        sends, recvs = {}, {}
        for r in range(world):
            for kind, peer, level in dt.rtree_steps(r, world):
                (sends if kind == "send" else recvs).setdefault((min(r, peer), max(r, peer), level), []).append(r)
        assert sends.keys() == recvs.keys()                      # every send has its matching recv at the same level
        assert len(sends) == world - 1                           # a tree: P - 1 messages
        assert all(len(dt.rtree_steps(r, world)) <= dt.tree_depth(world) for r in range(world))
        assert not any(k == "send" for k, _, _ in dt.rtree_steps(0, world))   # rank 0 ends up with R


# ---- CAQR schedule (cuda-qr_b200/dist_caqr.py) with numpy stand-ins for the GPU steps -----------------------
def _np_geqrf(a):
    """Unblocked Householder QR in place (LAPACK storage, reference sign convention qr.c:149-152); returns tau."""
    m, n = a.shape
    tau = np.zeros(n)
    for j in range(min(m, n)):
        x = a[j:, j].copy()
        nrm = np.linalg.norm(x)
        if nrm == 0.0:
            continue
        beta = nrm if x[0] < 0 else -nrm
        u = x[0] - beta
        v = x / u
        v[0] = 1.0
        tau[j] = -u / beta
        a[j:, j + 1:] -= tau[j] * np.outer(v, v @ a[j:, j + 1:])
        a[j, j] = beta
        a[j + 1:, j] = v[1:]
    return tau


def _np_apply_qt(a, tau, c):
    m, n = a.shape
    for j in range(min(m, n)):
        v = a[j:, j].copy()
        v[0] = 1.0
        c[j:] -= tau[j] * np.outer(v, v @ c[j:])


def _caqr_worker(rank, world, port, m_loc, n, kb, out):
    import sys
    sys.path.insert(0, ROOT)
    dc = importlib.import_module("cuda-qr_b200.dist_caqr")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    A = rng.standard_normal((m_loc * world, n))
    loc = A[rank * m_loc:(rank + 1) * m_loc].copy()
    taus, tree = {}, {}

    def local_qr(k0, w, r0):
        blk = loc[r0:, k0:k0 + w]
        taus[k0] = _np_geqrf(blk)

    def local_apply(k0, w, r0):
        _np_apply_qt(loc[r0:, k0:k0 + w], taus[k0], loc[r0:, k0 + w:])

    def gather(k0, w, r0):
        nt = n - (k0 + w)
        mine = np.hstack([np.triu(loc[r0:r0 + w, k0:k0 + w]), loc[r0:r0 + w, k0 + w:]])
        parts = [torch.empty((w, w + nt), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(mine)))
        full = np.vstack([p.numpy() for p in parts])
        return full[:, :w].copy(), (full[:, w:].copy() if nt else None)

    def tree_qr(k0, w, rs):
        tree[k0] = _np_geqrf(rs)

    def tree_apply(k0, w, rs, cs):
        _np_apply_qt(rs, tree[k0], cs)

    def put_back(k0, w, r0, rs, cs):
        top = loc[r0:r0 + w, k0:k0 + w]
        # the stacked-triangle structure must survive exactly: everything below the diagonal of my slice is zero
        assert np.all(np.tril(rs[rank * w:(rank + 1) * w], -1) == 0.0) or rank == 0
        top[:] = np.tril(top, -1) + np.triu(rs[rank * w:(rank + 1) * w])
        if cs is not None:
            loc[r0:r0 + w, k0 + w:] = cs[rank * w:(rank + 1) * w]

    dc.caqr_generic(rank, world, m_loc, n, kb, local_qr, local_apply, gather, tree_qr, tree_apply, put_back)
    if rank == 0:
        r = np.triu(loc[:n])
        ref = np.linalg.qr(A, mode="r")
        s1, s2 = np.sign(np.diag(r)), np.sign(np.diag(ref))
        out.put(float(np.linalg.norm(r * s1[:, None] - ref * s2[:, None]) / np.linalg.norm(ref)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,m_loc,n,kb", [(2, 40, 24, 8), (3, 30, 20, 8), (4, 64, 32, 16), (1, 50, 20, 8)])
def test_caqr_schedule_reduces_to_the_global_r(world, m_loc, n, kb):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_caqr_worker, args=(r, world, port, m_loc, n, kb, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=10) < 1e-12


def test_caqr_block_plan_and_shape_checks():
    dc = importlib.import_module("cuda-qr_b200.dist_caqr")
    assert dc.block_plan(100, 20, 8, 0) == [(0, 8, 0), (8, 8, 8), (16, 4, 16)]
    assert dc.block_plan(100, 20, 8, 3) == [(0, 8, 0), (8, 8, 0), (16, 4, 0)]
    with pytest.raises(ValueError):
        dc.check_shape(10, 20, 8, 2)
    with pytest.raises(ValueError):
        dc.check_shape(100, 20, 6, 2)
