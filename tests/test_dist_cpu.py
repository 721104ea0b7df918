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
