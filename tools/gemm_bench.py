"""Stand-alone timing of the tcgen05 3xTF32 GEMMs at trailing-update shapes (CUDA events, L2-cold operands > 126 MB)
and the per-class profile of one 16384^2 geqrf with look-ahead off (clean, non-overlapped class times).
    python tools/gemm_bench.py [gemm|geqrf] ..."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
what = sys.argv[1] if len(sys.argv) > 1 else "gemm"


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


if what == "gemm":
    once = len(sys.argv) > 2 and sys.argv[2] == "once"
    for (M, N, K) in [(16384, 16128, 256), (8192, 8192, 256), (16384, 16128, 64), (16384, 16128, 512)]:
        V = pkg.colmajor(M, K); V.copy_(torch.randn((M, K), device="cuda"))
        X = pkg.colmajor(K, N); X.copy_(torch.randn((K, N), device="cuda"))
        C = pkg.colmajor(M, N); C.copy_(torch.randn((M, N), device="cuda"))
        W = pkg.colmajor(K, N)
        if once:
            ctx.gemm_tf32x3(V, X, C, trans_a=False, alpha=-1.0, beta=1.0); ctx.gemm_tf32x3(V, C, W, trans_a=True); ctx.synchronize()
            break
        for beta in (0.0, 1.0):
            best, avg = timeit(lambda: ctx.gemm_tf32x3(V, X, C, trans_a=False, alpha=-1.0, beta=beta))
            fl = 2.0 * M * N * K
            print(f"NN  C{'-=' if beta else '='}V X  M={M} N={N} K={K} beta={beta}: best {best:.3f} ms avg {avg:.3f} ms  {fl / best / 1e9:.1f} TF/s (fp32-equivalent)  "
                  f"C traffic {(2 if beta else 1) * 4.0 * M * N / best / 1e6:.0f} GB/s", flush=True)
        best, avg = timeit(lambda: ctx.gemm_tf32x3(V, C, W, trans_a=True))
        print(f"TN  W = V^T C  M={K} N={N} K={M}: best {best:.3f} ms avg {avg:.3f} ms  {2.0 * M * N * K / best / 1e9:.1f} TF/s  C read {4.0 * M * N / best / 1e6:.0f} GB/s", flush=True)
        del V, X, C, W
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    ctx.set_option(pkg.OPT_LOOKAHEAD, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    A0 = pkg.colmajor(n, n); A0.copy_(torch.rand((n, n), device="cuda"))
    A = pkg.colmajor(n, n); tau = torch.zeros(n, device="cuda")
    for _ in range(2):
        A.copy_(A0); ctx.geqrf(A, tau)
    torch.cuda.synchronize()
    A.copy_(A0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.geqrf(A, tau); e1.record(); torch.cuda.synchronize()
    print(f"geqrf {n}^2 lookahead={ctx.get_option(pkg.OPT_LOOKAHEAD)}: {e0.elapsed_time(e1):.2f} ms")
    A.copy_(A0); torch.cuda.synchronize()
    ctx.profile_begin(); ctx.geqrf(A, tau); prof = ctx.profile_end()
    for k, v in prof.items():
        print(f"  {k:8s} {v['ms']:9.3f} ms  launches {v['launches']:6d}  {v['flops'] / max(v['ms'], 1e-9) / 1e9:9.1f} TF/s  {v['bytes'] / max(v['ms'], 1e-9) / 1e6:9.0f} GB/s (algorithmic)")
