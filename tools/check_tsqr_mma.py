"""R-only TSQR: tensor-pipe flat-tree leaf (OPT_FLAT_TSQR = 2) against the SIMT flat leaf and fp64 on a few shapes.
   python tools/check_tsqr_mma.py"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import metrics
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
g = torch.Generator(device="cuda").manual_seed(3)
ok = True
MODE = int(os.environ.get("CHECK_MODE", "2"))
for m, n, kind in [(16384, 64, "u"), (16448, 64, "n"), (20011, 17, "u"), (65536, 64, "u"), (100003, 40, "n"), (1 << 20, 64, "u"), (300000, 8, "n"), (77777, 63, "u")]:
    A = pkg.colmajor(m, n)
    A.copy_(torch.rand((m, n), device="cuda", generator=g) if kind == "u" else torch.randn((m, n), device="cuda", generator=g))
    G = A.t().double() @ A.double()
    out = []
    Rs = {}
    for mode in (MODE, 1):
        ctx.set_option(pkg.OPT_FLAT_TSQR, mode)
        R = pkg.colmajor(n, n); R.fill_(float("nan"))
        ctx.tsqr_r(A, R); ctx.synchronize()
        Rd = torch.triu(R.double())
        gram = float((Rd.t() @ Rd - G).norm() / G.norm())
        low = float(torch.tril(R, -1).abs().max()) if n > 1 else 0.0
        Rs[mode] = R.cpu().numpy()
        out.append(f"mode {mode}: gram {gram:.2e} below-diag {low:.1e}")
        ok = ok and gram < 5e-6 and low == 0.0
    R64 = np.linalg.qr(A.cpu().numpy().astype(np.float64), mode="r") if m <= 300000 else None
    d21 = metrics.r_rel_diff(Rs[MODE], Rs[1])
    d64 = metrics.r_rel_diff(Rs[MODE], R64) if R64 is not None else float("nan")
    ok = ok and d21 < 1e-4
    print(f"{m:8d} x {n:2d} {kind}: " + "  ".join(out) + f"  |R2-R1|/|R1| {d21:.2e}  |R2-R64| {d64:.2e}", flush=True)
print("TSQR_MMA_OK" if ok else "TSQR_MMA_FAIL")
