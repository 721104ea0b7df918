"""Diagnostics of the sliced Gram matrix (gram_umma.cu): compares the fp64 G the kernel reduced with the exact Gram matrix."""
import ctypes, importlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
lib = pkg.lib
lib.cqr_debug_gram_matrix.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
gen = torch.Generator(device="cuda").manual_seed(5)
ctx.set_option(pkg.OPT_FLAT_TSQR, 4)


def gram_of(X):
    A = pkg.colmajor(X.shape[0], X.shape[1]); A.copy_(X.float())
    R = pkg.colmajor(X.shape[1], X.shape[1])
    ctx.tsqr_r(A, R); ctx.synchronize()
    G = np.zeros((64, 64))
    rc = lib.cqr_debug_gram_matrix(ctx.h, G.ctypes.data)
    assert rc == 0, rc
    Gx = (A.double().t() @ A.double()).cpu().numpy()
    n = X.shape[1]
    G = G[:n, :n] + G[:n, :n].T     # the kernel keeps T = sum of the slabs' transposed halves; G = T + T^T
    global lastR
    lastR = np.triu(R.double().cpu().numpy())
    return G, Gx, ctx.tsqr_gram_info()


def report(name, X):
    G, Gx, info = gram_of(X)
    E = G - Gx
    Rk = np.linalg.cholesky(G).T
    Rx = np.linalg.cholesky(Gx).T
    print(f"   R: kernel vs chol(kernel G) {np.linalg.norm(lastR - Rk) / np.linalg.norm(Rk):.3e}   kernel vs chol(exact G) {np.linalg.norm(lastR - Rx) / np.linalg.norm(Rx):.3e}")
    print(f"{name:40s} |G-Gx|_F/|Gx|_F {np.linalg.norm(E) / np.linalg.norm(Gx):.3e}  max|E|/max|Gx| {np.abs(E).max() / np.abs(Gx).max():.3e}  "
          f"diag rel {np.abs(np.diag(E) / np.diag(Gx)).max():.3e}  info {info}", flush=True)
    return E, Gx


m = 16384
I8 = torch.randint(0, 256, (m, 64), device="cuda", generator=gen).double() / 256
report("8-bit fractions (s1 only)", I8)
I8s = torch.randint(-255, 256, (m, 64), device="cuda", generator=gen).double() / 256
report("signed 8-bit fractions (s1 only)", I8s)
I16 = torch.randint(0, 65536, (m, 64), device="cuda", generator=gen).double() / 65536
report("16-bit fractions (s1, s2)", I16)
I24 = torch.randint(0, 1 << 24, (m, 64), device="cuda", generator=gen).double() / (1 << 24)
E, Gx = report("24-bit fractions (s1, s2, s3)", I24)
print("   E[0:4,0:4] / Gx:", (E[:4, :4] / Gx[:4, :4]).round(10).tolist())
report("uniform fp32", torch.rand((m, 64), device="cuda", generator=gen))
report("normal fp32", torch.randn((m, 64), device="cuda", generator=gen))
for mm in (131072, 131072 + 128 * 5, 65536, 262144):
    report(f"uniform fp32, {mm} rows", torch.rand((mm, 64), device="cuda", generator=gen))
    report(f"normal fp32, {mm} rows", torch.randn((mm, 64), device="cuda", generator=gen))
report("uniform fp32, 1M rows", torch.rand((1 << 20, 64), device="cuda", generator=gen))
# one nonzero column pair: isolates single products
X = torch.zeros((m, 64), device="cuda"); X[:, 0] = 1.0; X[:, 1] = 1.0 / 65536 + 1.0 / 256
G, Gx, _ = gram_of(X)
print("two columns (1, 1/256 + 1/65536): G00 G01 G11 =", G[0, 0], G[0, 1], G[1, 1], " exact:", Gx[0, 0], Gx[0, 1], Gx[1, 1])
X = torch.zeros((m, 64), device="cuda"); X[:, 5] = 3.0; X[:, 40] = 1.0 + 2.0 ** -20
G, Gx, _ = gram_of(X)
print("two columns (3, 1 + 2^-20): G55 G5,40 G40,40 =", G[5, 5], G[5, 40], G[40, 40], " exact:", Gx[5, 5], Gx[5, 40], Gx[40, 40])
