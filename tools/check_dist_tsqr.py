"""Multi-GPU parity check (run under torchrun, one rank per GPU): row-partitioned TSQR with the NCCL
R-tree and the distributed thin Q, compared on rank 0 with the fp64 QR of the gathered matrix."""
import importlib, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import metrics
pkg = importlib.import_module("cuda-qr_b200")
dt = importlib.import_module("cuda-qr_b200.dist_tsqr")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ctx = pkg.Context(lr); ctx.use_torch_stream()
ok = True
for m_loc, n in [(1000, 64), (65536, 64), (5000, 48)]:
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    A = pkg.colmajor(m_loc, n, device=dev); A.copy_(torch.rand((m_loc, n), device=dev, generator=g))
    A0 = A.clone()
    ts = dt.DistTSQR(pkg, ctx, n, rank, world, dev)
    ts.factor(A, keep_q=True)
    Q = pkg.colmajor(m_loc, n, device=dev)
    ts.form_q(Q)
    ts.broadcast_r()
    torch.cuda.synchronize()
    gathered_A = [torch.empty((m_loc, n), device=dev) for _ in range(world)]
    gathered_Q = [torch.empty((m_loc, n), device=dev) for _ in range(world)]
    dist.all_gather(gathered_A, A0.contiguous()); dist.all_gather(gathered_Q, Q.contiguous())
    if rank == 0:
        Af = torch.cat(gathered_A).cpu().numpy(); Qf = torch.cat(gathered_Q).cpu().numpy(); R = ts.R.cpu().numpy()
        be, orth = metrics.backward_error(Af, Qf, np.triu(R)), metrics.orthogonality(Qf)
        dr = metrics.r_rel_diff(R, np.linalg.qr(Af.astype(np.float64), mode="r"))
        print(f"world {world} m_loc {m_loc} n {n}: backward {be:.3f} orth {orth:.3f} dR {dr:.2e}")
        ok = ok and be <= 10 and orth <= 10 and dr <= 1e-4
# R-only path (flat-tree warp leaf on every rank) at a size where it is active: Gram check against all ranks' rows
for m_loc, n in [(262144, 64), (100003, 40)]:
    g = torch.Generator(device=dev).manual_seed(50 + rank)
    A = pkg.colmajor(m_loc, n, device=dev); A.copy_(torch.rand((m_loc, n), device=dev, generator=g))
    ts = dt.DistTSQR(pkg, ctx, n, rank, world, dev)
    ts.factor(A, keep_q=False)
    G = A.t().double() @ A.double()
    dist.all_reduce(G)
    torch.cuda.synchronize()
    if rank == 0:
        Rd = torch.triu(ts.R.double())
        ge = float((Rd.t() @ Rd - G).norm() / G.norm())
        print(f"world {world} m_loc {m_loc} n {n}: R-only gram error {ge:.2e}")
        ok = ok and ge < 1e-5
dist.barrier(); dist.destroy_process_group()
if rank == 0:
    print("DIST_TSQR_OK" if ok else "DIST_TSQR_FAIL")
    sys.exit(0 if ok else 1)
