"""Row-partitioned R-only TSQR across the ranks of a torchrun launch: the peer-memory R tree (cqr_tsqr_dist_r) against the
ncclSend / ncclRecv tree, CUDA events, max over ranks, Gram check of the combined R.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_tsqr_bench.py [rows]"""
import importlib, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG", "WARN")
pkg = importlib.import_module("cuda-qr_b200")
dt = importlib.import_module("cuda-qr_b200.dist_tsqr")
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = pkg.Context(lr); ctx.use_torch_stream()
rows = [int(a) for a in sys.argv[1:]] or [8388608]
n = 64


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in ev:
        e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


for m_total in rows:
    m_loc = m_total // world
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    A = pkg.colmajor(m_loc, n, device=dev); A.copy_(torch.rand((m_loc, n), device=dev, generator=g))
    G = A.t().double() @ A.double()
    if world > 1:
        dist.all_reduce(G)
    R_loc = pkg.colmajor(n, n, device=dev)
    t_local = timed(lambda: ctx.tsqr_r(A, R_loc))
    out = [f"local TSQR {t_local:.3f} ms"]
    for mode in (["nccl", "peer"] if world > 1 else ["single"]):
        ts = dt.DistTSQR(pkg, ctx, n, rank, world, dev)
        if mode == "peer":
            ts.enable_peer()
        t = timed(lambda: ts.factor(A, keep_q=False))
        gram = float("nan")
        if rank == 0:
            Rd = torch.triu(ts.R.double())
            gram = float((Rd.t() @ Rd - G).norm() / G.norm())
        out.append(f"{mode} tree: {t:.3f} ms (cross-GPU part {t - t_local:+.3f} ms, gram {gram:.2e})")
    if rank == 0:
        print(f"world {world}  {m_total} x {n}: " + "   ".join(out), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
