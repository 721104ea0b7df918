"""cqr_geqrf over a range of shapes (median of 5, CUDA events) -- run it under different knob settings to check that a
default tuned at 16384^2 does not cost elsewhere.   python tools/size_sweep.py"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
shapes = [(1024, 1024), (2048, 2048), (4096, 4096), (8192, 8192), (12288, 12288), (16384, 4096), (16384, 16384), (20480, 8192), (32768, 4096)]
for m, n in shapes:
    A0 = pkg.colmajor(m, n); A0.copy_(torch.rand((m, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(12)))
    A = pkg.colmajor(m, n); tau = torch.zeros(n, device="cuda")
    for _ in range(2):
        A.copy_(A0); ctx.geqrf(A, tau)
    ts = []
    for _ in range(5):
        A.copy_(A0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.geqrf(A, tau); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    Rd = torch.triu(A[:n].double()); G = A0.t().double() @ A0.double()
    fl = 2.0 * m * n * n - 2.0 * n ** 3 / 3
    print(f"geqrf {m:6d} x {n:5d}: {ts[2]:8.3f} ms  {fl / ts[2] / 1e9:6.1f} TF/s  gram {float((Rd.t() @ Rd - G).norm() / G.norm()):.1e}", flush=True)
    del A0, A, Rd, G
