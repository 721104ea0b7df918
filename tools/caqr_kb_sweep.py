"""CAQR (config 5 per-rank shape, 16384 x 4096 per GPU) for several outer block widths kb, under torchrun.
   python -m torch.distributed.run --nproc-per-node N ... tools/caqr_kb_sweep.py [kb ...]"""
import importlib, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG", "WARN")
pkg = importlib.import_module("cuda-qr_b200")
dc = importlib.import_module("cuda-qr_b200.dist_caqr")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ctx = pkg.Context(lr); ctx.use_torch_stream()
m_loc, n = 16384, 4096
g = torch.Generator(device=dev).manual_seed(300 + rank)
A0 = pkg.colmajor(m_loc, n, device=dev); A0.copy_(torch.rand((m_loc, n), device=dev, generator=g))
A = pkg.colmajor(m_loc, n, device=dev)
G = A0.t().double() @ A0.double(); dist.all_reduce(G)
for kb in [int(a) for a in sys.argv[1:]] or [256, 512, 1024]:
    cq = dc.DistCAQR(pkg, ctx, m_loc, n, rank, world, dev, kb=kb)
    ts = []
    for it in range(4):
        A.copy_(A0); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cq.factor(A); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ts.append(float(t))
    if rank == 0:
        R = pkg.colmajor(n, n, device=dev); cq.extract_r(A, R); Rd = torch.triu(R.double())
        gram = float((Rd.t() @ Rd - G).norm() / G.norm())
        fl = 2.0 * m_loc * world * n * n - 2.0 * n ** 3 / 3
        print(f"world {world} kb {kb}: {min(ts[1:]):.2f} ms  {fl / min(ts[1:]) / 1e9:.1f} TFLOP/s aggregate  gram {gram:.2e}", flush=True)
    del cq
dist.barrier(); dist.destroy_process_group()
