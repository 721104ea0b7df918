"""Gram check of cqr_geqrf for a given outer block width (argv: n outer [reps]) -- regression probe."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
n = int(sys.argv[1]); outer = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx.set_option(pkg.OPT_OUTER_BLOCK, outer)
A0 = pkg.colmajor(n, n); A0.copy_(torch.rand((n, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(12)))
G = A0.t().double() @ A0.double()
A = pkg.colmajor(n, n); tau = torch.zeros(n, device="cuda")
for r in range(reps):
    A.copy_(A0); ctx.geqrf(A, tau); rc = ctx.synchronize() if hasattr(ctx, "synchronize") else 0
    Rd = torch.triu(A[:n].double())
    E = (Rd.t() @ Rd - G)
    colerr = E.norm(dim=0) / G.norm(dim=0)
    bad = torch.nonzero(colerr > 1e-3).flatten()
    print(f"n={n} outer={outer} rep {r}: gram {float(E.norm() / G.norm()):.2e}  first bad column {int(bad[0]) if len(bad) else -1}  bad columns {len(bad)}", flush=True)
