"""Batched 64x64 QR timing (cqr_geqrf_batched), CUDA events, median of 15; CQR_BATCHED_CTA=1 selects the CTA kernel."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
once = len(sys.argv) > 2
A0 = torch.rand((batch, 64, 64), device="cuda", generator=torch.Generator(device="cuda").manual_seed(12))
A = torch.empty_like(A0); tau = torch.zeros((batch, 64), device="cuda")
if once:
    A.copy_(A0); ctx.geqrf_batched(A, tau); ctx.synchronize(); sys.exit(0)
for _ in range(3):
    A.copy_(A0); ctx.geqrf_batched(A, tau)
ctx.synchronize()
ts = []
for _ in range(15):
    A.copy_(A0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.geqrf_batched(A, tau); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
Rd = torch.triu(A.transpose(1, 2).double())
G = A0.double() @ A0.double().transpose(1, 2)
err = float(((Rd.transpose(1, 2) @ Rd - G).flatten(1).norm(dim=1) / G.flatten(1).norm(dim=1)).max())
flops = batch * (2.0 * 64 ** 3 - 2.0 * 64 ** 3 / 3)
print(f"batched {batch} x 64x64 {'cta' if os.environ.get('CQR_BATCHED_CTA') else 'warp'}: median {ts[7]:.3f} ms min {ts[0]:.3f} ms  {flops / ts[7] / 1e9:.1f} TFLOP/s  "
      f"{2.0 * batch * 64 * 64 * 4 / ts[7] / 1e6:.1f} GB/s  max gram err {err:.2e}", flush=True)
