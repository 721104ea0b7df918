"""Gram leaf of the R-only TSQR (gram_umma.cu, OPT_FLAT_TSQR = 4): accuracy against fp64, exactness of the sliced Gram
matrix, the condition gate, timing against the Householder leaf.
   python tools/gram_check.py [quick]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(12)


def ref_r(A):
    """fp64 R (positive diagonal) of A by Cholesky of the fp64 Gram matrix when well conditioned, else Householder."""
    Ad = A.double()
    m = Ad.shape[0]
    if m <= 1 << 20:
        R = torch.linalg.qr(Ad, mode="r").R
    else:   # blocked TSQR in fp64: stack the R factors of 1M-row pieces
        Rs = [torch.linalg.qr(Ad[i:i + (1 << 20)], mode="r").R for i in range(0, m, 1 << 20)]
        R = torch.linalg.qr(torch.cat(Rs), mode="r").R
    sgn = torch.sign(torch.diagonal(R)); sgn[sgn == 0] = 1
    return R * sgn[:, None]


def norm_sign(R):
    R = torch.triu(R.double())
    sgn = torch.sign(torch.diagonal(R)); sgn[sgn == 0] = 1
    return R * sgn[:, None]


def run(A, leaf):
    m, n = A.shape
    R = pkg.colmajor(n, n)
    ctx.set_option(pkg.OPT_FLAT_TSQR, leaf)
    ctx.tsqr_r(A, R); ctx.synchronize()
    info = (0.0, True)
    if leaf == 4:
        try:
            info = ctx.tsqr_gram_info()
        except pkg.CudaQRError:      # shape not eligible for the Gram leaf (lda % 4 != 0): the Householder leaf ran
            info = (float("nan"), True)
    return norm_sign(R), info


def timeit(A, leaf, reps=15):
    m, n = A.shape
    R = pkg.colmajor(n, n)
    ctx.set_option(pkg.OPT_FLAT_TSQR, leaf)
    for _ in range(3):
        ctx.tsqr_r(A, R)
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.tsqr_r(A, R); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def case(name, A):
    m, n = A.shape
    Rref = ref_r(A)
    out = f"{name:34s} {m:8d} x {n:2d}:"
    for leaf in (4, 1):
        R, (bound, hh) = run(A, leaf)
        err = float((R - Rref).norm() / Rref.norm())
        G = A.t().double() @ A.double()
        gram = float((R.t() @ R - G).norm() / G.norm())
        tag = "gram" if leaf == 4 else "hh  "
        out += f"  {tag} |R-R64|/|R64| {err:.2e} gramerr {gram:.2e}"
        if leaf == 4:
            out += f" bound {bound:.3g} fallback {int(hh)}"
    print(out, flush=True)


def graded(m, n, cond):
    """random matrix with singular values from 1 down to 1/cond (geometric)"""
    Q = torch.linalg.qr(torch.randn((m, n), device=dev, generator=gen, dtype=torch.float64)).Q
    V = torch.linalg.qr(torch.randn((n, n), device=dev, generator=gen, dtype=torch.float64)).Q
    s = torch.logspace(0, -torch.log10(torch.tensor(float(cond))).item(), n, device=dev, dtype=torch.float64)
    return (Q * s) @ V.t()


def cm(X):
    A = pkg.colmajor(X.shape[0], X.shape[1]); A.copy_(X.float()); return A


rows = 131072 if quick else 1048576
case("uniform[0,1)", cm(torch.rand((rows, 64), device=dev, generator=gen)))
case("normal", cm(torch.randn((rows, 64), device=dev, generator=gen)))
case("uniform ragged m, n = 40", cm(torch.rand((rows - 76, 40), device=dev, generator=gen)))
case("uniform, lda % 4 != 0 (not eligible)", cm(torch.rand((rows - 77, 64), device=dev, generator=gen)))
case("column scales 2^-20 .. 2^20", cm(torch.randn((rows, 64), device=dev, generator=gen) * (2.0 ** torch.linspace(-20, 20, 64, device=dev))))
case("rows graded 1 .. 1e-6", cm(torch.randn((rows, 64), device=dev, generator=gen) * torch.logspace(0, -6, rows, device=dev)[:, None]))
for cond in (10, 30, 100, 300, 1000, 1e4):
    case(f"singular values 1 .. 1/{cond:g}", cm(graded(131072, 64, cond)))
X = torch.rand((131072, 64), device=dev, generator=gen); X[:, 7] = X[:, 3]
case("duplicated column (singular)", cm(X))
X = torch.rand((131072, 64), device=dev, generator=gen); X[:, 9] = 0
case("zero column", cm(X))
X = torch.rand((131072, 64), device=dev, generator=gen); X[5, 5] = float("inf")
try:
    case("Inf entry", cm(X))
except Exception as e:   # fp64 reference may refuse
    print("Inf entry: reference failed:", type(e).__name__)
X = torch.rand((131072, 64), device=dev, generator=gen) * 1e-30
case("all entries ~1e-30 (out of scale)", cm(X))

# exactness: integer-valued data whose Gram matrix is exactly representable -> the sliced Gram must reproduce it to the last bit
Ai = torch.randint(-2000, 2001, (131072, 64), device=dev, generator=gen).float()
R, (bound, hh) = run(cm(Ai), 4)
Gd = Ai.t().double() @ Ai.double()
print(f"integer data: |R^T R - G|/|G| = {float((R.t() @ R - Gd).norm() / Gd.norm()):.2e} (fp32 rounding of R only), bound {bound:.3g}", flush=True)

for m in ([131072, 1048576] if quick else [8388608, 1048576, 131072, 16384]):
    A = cm(torch.rand((m, 64), device=dev, generator=gen))
    tg, tgmin = timeit(A, 4)
    th, thmin = timeit(A, 1)
    byts = 4.0 * m * 64
    print(f"time {m:8d} x 64: gram leaf median {tg:7.3f} ms (min {tgmin:7.3f}) = {byts / tg / 1e6:7.1f} GB/s   householder leaf {th:7.3f} ms (min {thmin:7.3f}) = {byts / th / 1e6:7.1f} GB/s",
          flush=True)
    del A
ctx.set_option(pkg.OPT_FLAT_TSQR, 1)
