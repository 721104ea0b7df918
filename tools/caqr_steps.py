"""Per-step CUDA-event brackets of one DistCAQR.factor (config 5 per-rank shape) under torchrun: where a block's time goes.
   python -m torch.distributed.run --nproc-per-node N ... tools/caqr_steps.py"""
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
A0 = pkg.colmajor(m_loc, n, device=dev); A0.copy_(torch.rand((m_loc, n), device=dev, generator=torch.Generator(device=dev).manual_seed(300 + rank)))
A = pkg.colmajor(m_loc, n, device=dev)
cq = dc.DistCAQR(pkg, ctx, m_loc, n, rank, world, dev)
for it in range(3):
    cq.profile = it == 2
    A.copy_(A0); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cq.factor(A); e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"factor {e0.elapsed_time(e1):.2f} ms" + (f"  steps {dict((k, round(v, 2)) for k, v in cq.step_ms.items())}" if cq.profile else ""), flush=True)
dist.barrier(); dist.destroy_process_group()
