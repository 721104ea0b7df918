"""One device-resident cqr_geqrf of an n x n matrix and nothing else: the ncu launch-list target.
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python tools/one_geqrf.py [n]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
A = pkg.colmajor(n, n); A.copy_(torch.rand((n, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(12)))
tau = torch.zeros(n, device="cuda")
ctx.geqrf(A, tau); ctx.synchronize()
print("launches", ctx.launch_count())
