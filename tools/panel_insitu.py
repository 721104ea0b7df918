"""In-situ panel durations of one square geqrf (library event brackets of class "panel", in launch order): mean per group
of 32 panels, so the effect of a panel-kernel change inside the real pipeline can be compared with tools/panel_bench.py.
    CQR_PANEL_PAIR=0 python tools/panel_insitu.py [n]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
A0 = pkg.colmajor(n, n); A0.copy_(torch.rand((n, n), device="cuda"))
A = pkg.colmajor(n, n); tau = torch.zeros(n, device="cuda")
for _ in range(2):
    A.copy_(A0); ctx.geqrf(A, tau)
torch.cuda.synchronize()
A.copy_(A0); torch.cuda.synchronize()
ctx.profile_begin(); ctx.geqrf(A, tau)
tl = ctx.profile_timeline()
ctx.profile_end()
pan = [(t1 - t0) * 1e3 for t0, t1, c in sorted(tl) if c == "panel"]
end = max(t[1] for t in tl)
print(f"CQR_PANEL_PAIR={os.environ.get('CQR_PANEL_PAIR', '(default)')} n={n}: {len(pan)} panels, sum {sum(pan) / 1e3:.2f} ms, last bracket ends {end:.2f} ms")
g = 32
for i in range(0, len(pan), g):
    seg = pan[i:i + g]
    print(f"  panels {i:3d}..{i + len(seg) - 1:3d} (rows {n - 64 * i:5d} ..): mean {sum(seg) / len(seg):6.1f} us  min {min(seg):6.1f}  max {max(seg):6.1f}")
