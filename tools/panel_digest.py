"""Digest of cqr_geqrf results on m x 64 panels (A and tau, sha1 of the raw bytes): run under two builds of the library
(CQR_LIB=...) to check that a kernel refactoring left the arithmetic bit-identical.   python tools/panel_digest.py [m ...]"""
import hashlib, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
for m in [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192, 16384]:
    for kind in ("uniform", "normal", "dependent"):
        g = torch.Generator(device="cuda").manual_seed(m + 7)
        X = torch.rand((m, 64), device="cuda", generator=g) if kind == "uniform" else torch.randn((m, 64), device="cuda", generator=g)
        if kind == "dependent":                       # neighbouring columns nearly parallel: the pair guard must fall back
            X[:, 1::2] = X[:, 0::2] + 1e-3 * X[:, 1::2]
        A = pkg.to_colmajor(X); tau = torch.zeros(64, device="cuda")
        ctx.geqrf(A, tau); ctx.synchronize()
        ha = hashlib.sha1(A.t().contiguous().cpu().numpy().tobytes()).hexdigest()[:16]
        ht = hashlib.sha1(tau.cpu().numpy().tobytes()).hexdigest()[:16]
        R = torch.triu(A[:64].double()); G = X.double().t() @ X.double()
        print(f"m={m:6d} {kind:9s} A {ha} tau {ht}  gram {float((R.t() @ R - G).norm() / G.norm()):.2e}", flush=True)
