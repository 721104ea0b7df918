"""geqrf timing for the look-ahead schemes (CQR_OPT_LOOKAHEAD 1: one K = 256 slice per block, 2: panel-wise slices in the
panel-bound phase): median of 7, CUDA events; also checks that both give the same R."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
if os.environ.get("LA_OUTER"):
    ctx.set_option(pkg.OPT_OUTER_BLOCK, int(os.environ["LA_OUTER"]))
for n in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192, 16384]:
    A0 = pkg.colmajor(n, n); A0.copy_(torch.rand((n, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(12)))
    A = pkg.colmajor(n, n); tau = torch.zeros(n, device="cuda")
    out = []
    ref = None
    for la in (1, 2):
        ctx.set_option(pkg.OPT_LOOKAHEAD, la)
        for _ in range(2):
            A.copy_(A0); ctx.geqrf(A, tau)
        ts = []
        worst = 0.0
        for _ in range(7):
            A.copy_(A0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.geqrf(A, tau); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            R = torch.triu(A[:n])
            if ref is None: ref = R.clone()
            worst = max(worst, float((R - ref).norm() / ref.norm()))
        ts.sort()
        fl = 4.0 * n ** 3 / 3
        out.append(f"la={la}: {ts[3]:8.3f} ms {fl / ts[3] / 1e9:6.1f} TF/s (max |R-R1|/|R1| over 7 runs {worst:.1e})")
    print(f"geqrf {n:5d}^2  " + "   ".join(out), flush=True)
ctx.set_option(pkg.OPT_LOOKAHEAD, 1)
