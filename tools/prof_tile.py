"""Small driver for ncu captures: one TSQR (R-only) of 1M x 64, one batched 64x64 x 16384, one 4096^2 geqrf."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
what = sys.argv[1] if len(sys.argv) > 1 else "tsqr"
if what == "tsqr":
    m, n = 1 << 20, 64
    A = pkg.colmajor(m, n); A.copy_(torch.rand((m, n), device="cuda"))
    R = pkg.colmajor(n, n)
    for _ in range(2):
        ctx.tsqr_r(A, R)
    ctx.synchronize()
elif what == "batched":
    A = torch.rand((16384, 64, 64), device="cuda"); tau = torch.zeros((16384, 64), device="cuda")
    for _ in range(2):
        ctx.geqrf_batched(A, tau)
    ctx.synchronize()
else:
    m = n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    A0 = pkg.colmajor(m, n); A0.copy_(torch.rand((m, n), device="cuda"))
    A = pkg.colmajor(m, n); tau = torch.zeros(n, device="cuda")
    for _ in range(2):
        A.copy_(A0); ctx.geqrf(A, tau)
    ctx.synchronize()
