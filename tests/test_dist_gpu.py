"""Multi-rank parity on ONE device: world = 2 and 4 processes share cuda:0 and talk over gloo (messages staged through
the host, cuda-qr_b200/comm.py), so the box that runs `pytest -m gpu` exercises the real CUDA leaf / combine / apply
kernels of the row-partitioned TSQR (cuda-qr_b200/dist_tsqr.py) and CAQR (cuda-qr_b200/dist_caqr.py) paths -- only the
wire differs from the NCCL launch.  The reference has no multi-GPU code (qr.cu:737): parity is judged against the
oracle (`port.mmqr`, the restated qr.c) on a shape legal for it, against the fp64 QR, and against the single-GPU result
of the same matrix; the A = Q R contract is qr.c:330-438."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    pkg = importlib.import_module("cuda-qr_b200")
    ctx = pkg.Context(0)
    ctx.use_torch_stream()
    return pkg, ctx, torch.device("cuda", 0)


def _matrix(m, n, seed, kind="uniform"):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.random((m, n), dtype=np.float32)
    return rng.standard_normal((m, n)).astype(np.float32)


def _tsqr_worker(rank, world, port, m, n, keep_q, out):
    pkg, ctx, dev = _init(rank, world, port)
    dt = importlib.import_module("cuda-qr_b200.dist_tsqr")
    A = _matrix(m, n, 21)
    lo, hi = rank * m // world, (rank + 1) * m // world
    Al = pkg.to_colmajor(torch.from_numpy(np.ascontiguousarray(A[lo:hi])).to(dev))
    ts = dt.DistTSQR(pkg, ctx, n, rank, world, dev)
    ts.factor(Al, keep_q=keep_q)
    Q = None
    if keep_q:
        Ql = pkg.colmajor(hi - lo, n, device=dev)
        ts.form_q(Ql)
        Q = Ql.cpu().numpy()
    ctx.synchronize()
    out.put((rank, Q, ts.R.cpu().numpy() if rank == 0 else None, ctx.launch_count()))
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


def _run(worker, world, args, timeout=300):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port) + tuple(args) + (out,)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=timeout) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return sorted(res, key=lambda r: r[0])


@pytest.mark.parametrize("world", [2, 4])
def test_dist_tsqr_thin_q_and_r_match_oracle(world, port):
    """2044 x 64 is legal for the reference's window sweep (PR = 64, PC = 4: m = 64 + 60 k): R against the restated
    qr.c and fp64, thin Q through the tree against the acceptance bounds."""
    from oracle import metrics
    m, n = 2044, 64
    res = _run(_tsqr_worker, world, (m, n, True))
    A = _matrix(m, n, 21)
    Q = np.vstack([r[1] for r in res])
    R = np.triu(res[0][2])
    assert all(r[3] > 0 for r in res)                           # every rank launched CUDA kernels
    rv_ref, _ = port.mmqr(A, 64, 4)
    assert metrics.r_rel_diff(R, rv_ref) <= metrics.TOL_R
    assert metrics.r_rel_diff(R, np.linalg.qr(A.astype(np.float64), mode="r")) <= metrics.TOL_R
    assert metrics.backward_error(A, Q, R) <= metrics.TOL_BACKWARD
    assert metrics.orthogonality(Q) <= metrics.TOL_ORTH


@pytest.mark.parametrize("world,keep_q", [(2, False), (4, False), (4, True)])
def test_dist_tsqr_flat_leaf_ranks(world, keep_q):
    """Per-rank slabs tall enough for the warp-resident flat leaf (>= 16384 rows), ragged split."""
    from oracle import metrics
    m, n = 4 * 20011, 64
    res = _run(_tsqr_worker, world, (m, n, keep_q))
    A = _matrix(m, n, 21)
    R = np.triu(res[0][2])
    assert metrics.r_rel_diff(R, np.linalg.qr(A.astype(np.float64), mode="r")) <= metrics.TOL_R
    if keep_q:
        Q = np.vstack([r[1] for r in res])
        assert metrics.backward_error(A, Q, R) <= metrics.TOL_BACKWARD
        assert metrics.orthogonality(Q) <= metrics.TOL_ORTH


def _tsqr_peer_worker(rank, world, port, m, n, reps, out):
    pkg, ctx, dev = _init(rank, world, port)
    dt = importlib.import_module("cuda-qr_b200.dist_tsqr")
    ts = dt.DistTSQR(pkg, ctx, n, rank, world, dev).enable_peer()
    Rs = []
    for rep in range(reps):                                     # back-to-back calls: slots are reused with no host sync in between
        A = _matrix(m, n, 40 + rep)
        lo, hi = rank * m // world, (rank + 1) * m // world
        Al = pkg.to_colmajor(torch.from_numpy(np.ascontiguousarray(A[lo:hi])).to(dev))
        ts.factor(Al, keep_q=False)
        if rank == 0:
            Rs.append(ts.R.clone())
    ctx.synchronize()
    out.put((rank, [r.cpu().numpy() for r in Rs], ctx.launch_count()))
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


@pytest.mark.parametrize("world,m,n", [(2, 2044, 64), (4, 4 * 20011, 64), (3, 9000, 40), (4, 2044, 64), (5, 10240, 32), (8, 20000, 64)])
def test_dist_tsqr_peer_memory_rtree(world, m, n, port):
    """cqr_tsqr_dist_r: the cross-rank R combine as ONE kernel per rank over cudaIpc-mapped peer slabs (stores into the
    receiver's slab, flag release / acquire, stacked QR in registers; up to 8 ranks one hop into rank 0, which factors the
    stacked (64 world) x 64 matrix: tile_qr_core<4 / 8 / 16>), five calls back to back.  R against fp64, and on
    the reference-legal shape against the restated qr.c."""
    from oracle import metrics
    reps = 5
    res = _run(_tsqr_peer_worker, world, (m, n, reps))
    for rep in range(reps):
        A = _matrix(m, n, 40 + rep)
        R = np.triu(res[0][1][rep])
        assert metrics.r_rel_diff(R, np.linalg.qr(A.astype(np.float64), mode="r")) <= metrics.TOL_R
        if (m - 64) % 60 == 0 and n % 4 == 0 and rep == 0:
            rv_ref, _ = port.mmqr(A, 64, 4)
            assert metrics.r_rel_diff(R, rv_ref) <= metrics.TOL_R


def _caqr_worker(rank, world, port, m_loc, n, kb, kind, out):
    pkg, ctx, dev = _init(rank, world, port)
    dc = importlib.import_module("cuda-qr_b200.dist_caqr")
    A = _matrix(m_loc * world, n, 33, kind)
    Al = pkg.to_colmajor(torch.from_numpy(np.ascontiguousarray(A[rank * m_loc:(rank + 1) * m_loc])).to(dev))
    cq = dc.DistCAQR(pkg, ctx, m_loc, n, rank, world, dev, kb=kb)
    cq.factor(Al)
    R = pkg.colmajor(n, n, device=dev)
    cq.extract_r(Al, R)
    Ql = pkg.colmajor(m_loc, n, device=dev)
    cq.form_q(Al, Ql)
    # Q^T applied to the original rows must give [R; 0] (apply_q with trans), checked on the gathered result
    Cl = pkg.to_colmajor(torch.from_numpy(np.ascontiguousarray(A[rank * m_loc:(rank + 1) * m_loc])).to(dev))
    cq.apply_q(Al, Cl, True)
    ctx.synchronize()
    out.put((rank, Ql.cpu().numpy(), R.cpu().numpy() if rank == 0 else None, Cl.cpu().numpy(), ctx.launch_count()))
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


@pytest.mark.parametrize("world,m_loc,n,kb,kind", [(2, 1024, 512, 256, "uniform"), (4, 1024, 512, 256, "normal"),
                                                   (2, 300, 96, 32, "normal"), (4, 2048, 1024, 256, "uniform")])
def test_dist_caqr_r_and_q(world, m_loc, n, kb, kind, pkg):
    """R against the single-GPU cqr_geqrf R of the same matrix (<= 1e-4) and against fp64; thin Q through the local and
    tree reflectors against the acceptance bounds; Q^T A = [R; 0]."""
    from oracle import metrics
    res = _run(_caqr_worker, world, (m_loc, n, kb, kind))
    A = _matrix(m_loc * world, n, 33, kind)
    R = np.triu(res[0][2])
    Q = np.vstack([r[1] for r in res])
    QtA = np.vstack([r[3] for r in res])
    assert all(r[4] > 0 for r in res)
    ctx = pkg.Context(0)
    ctx.use_torch_stream()
    dA = pkg.to_colmajor(torch.from_numpy(A).cuda())
    tau = torch.zeros(n, device="cuda")
    ctx.geqrf(dA, tau)
    ctx.synchronize()
    R1 = np.triu(dA[:n].cpu().numpy())
    ctx.close()
    assert metrics.r_rel_diff(R, R1) <= metrics.TOL_R
    assert metrics.r_rel_diff(R, np.linalg.qr(A.astype(np.float64), mode="r")) <= metrics.TOL_R
    assert metrics.backward_error(A, Q, R) <= metrics.TOL_BACKWARD
    assert metrics.orthogonality(Q) <= metrics.TOL_ORTH
    # rank 0's first n rows of Q^T A are R; everything else is zero at the backward-error level
    top = QtA[:n]
    rest = np.vstack([QtA[n:m_loc]] + [QtA[p * m_loc:(p + 1) * m_loc] for p in range(1, world)])
    nrm = np.linalg.norm(A.astype(np.float64))
    assert np.linalg.norm(np.triu(top).astype(np.float64) - R) / (nrm * n * metrics.EPS32) <= metrics.TOL_BACKWARD
    assert np.sqrt(np.linalg.norm(np.tril(top, -1)) ** 2 + np.linalg.norm(rest) ** 2) / (nrm * n * metrics.EPS32) <= metrics.TOL_BACKWARD
