"""GPU accuracy probe: backward error / orthogonality of cqr_geqrf across sizes, GEMM modes and input
distributions, next to cuSOLVER's geqrf (torch.linalg.qr) as an external fp32 yardstick.
    python tools/accuracy_probe.py [sizes...]
Prints one line per case:  ||A-QR||_F/||A||_F, ||Q^T Q - I||_F/sqrt(n), both also in units of n*eps."""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
EPS = 2.0 ** -23
sizes = [int(s) for s in sys.argv[1:]] or [1024, 2048, 4096, 8192]
SIMT_MAX = int(os.environ.get("PROBE_SIMT_MAX", "16384"))                      # fp32 SIMT GEMMs up to this size
OUTERS = [int(x) for x in os.environ.get("PROBE_OUTERS", "256,64").split(",")]   # outer block widths


def blocked_norm_diff(X, Y):
    num = torch.zeros((), device="cuda", dtype=torch.float64); den = torch.zeros((), device="cuda", dtype=torch.float64)
    for c0 in range(0, X.shape[1], 2048):
        d = X[:, c0:c0 + 2048].double() - Y[:, c0:c0 + 2048].double()
        num += (d * d).sum(); den += (X[:, c0:c0 + 2048].double() ** 2).sum()
    return float((num / den).sqrt())


def ours(A0, mode, outer=256):
    m, n = A0.shape
    ctx.set_option(pkg.OPT_GEMM, mode)
    ctx.set_option(pkg.OPT_OUTER_BLOCK, outer)
    A = pkg.colmajor(m, n); A.copy_(A0)
    tau = torch.zeros(n, device="cuda")
    ctx.geqrf(A, tau); ctx.synchronize()
    R = pkg.colmajor(n, n); ctx.extract_r(A, R)
    QR = pkg.colmajor(m, n); QR.zero_(); QR[:n].copy_(R)
    ctx.apply_q(A, tau, QR, trans=False); ctx.synchronize()
    be = blocked_norm_diff(A0, QR)
    del QR
    Q = pkg.colmajor(m, n); ctx.form_q(A, tau, Q); ctx.synchronize()
    G = (Q.t().double() @ Q.double()) if n <= 16384 else None
    orth = float((G - torch.eye(n, device="cuda", dtype=torch.float64)).norm() / n ** 0.5) if G is not None else float("nan")
    return be, orth


def cusolver(A0):
    m, n = A0.shape
    t0 = time.perf_counter(); Q, R = torch.linalg.qr(A0.contiguous()); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    be = blocked_norm_diff(A0, Q @ R)
    G = Q.t().double() @ Q.double() if n <= 16384 else None
    orth = float((G - torch.eye(n, device="cuda", dtype=torch.float64)).norm() / n ** 0.5) if G is not None else float("nan")
    return be, orth, dt


for n in sizes:
    for dist in ("uniform", "normal"):
        g = torch.Generator(device="cuda").manual_seed(12)
        X = torch.rand((n, n), device="cuda", generator=g) if dist == "uniform" else torch.randn((n, n), device="cuda", generator=g)
        A0 = pkg.to_colmajor(X)
        for mode, name in ((1, "tf32x3"), (0, "simt")):
            if mode == 0 and n > SIMT_MAX:
                continue
            for outer in OUTERS:
                be, orth = ours(A0, mode, outer)
                print(f"n={n:6d} {dist:8s} ours/{name:7s} outer={outer:3d}: backward {be:.3e} ({be / (n * EPS):.4f} n*eps)  orth {orth:.3e} ({orth / (n * EPS):.4f} n*eps)", flush=True)
        if n <= 16384:
            be, orth, dt = cusolver(X)
            print(f"n={n:6d} {dist:8s} cusolver(torch.linalg.qr) {dt*1e3:8.1f} ms incl. Q: backward {be:.3e}  orth {orth:.3e}", flush=True)
ctx.set_option(pkg.OPT_GEMM, 1); ctx.set_option(pkg.OPT_OUTER_BLOCK, 256)
