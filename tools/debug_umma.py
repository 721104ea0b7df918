"""GPU debugging aid: exercise the tcgen05 3xTF32 GEMMs directly and print diagnostics."""
import importlib, sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
torch.manual_seed(0)

def run(trans, M, N, K, pattern="rand"):
    if pattern == "rand":
        A = torch.randn((K, M) if trans else (M, K), device="cuda")
        B = torch.randn((K, N), device="cuda")
    else:
        A = torch.zeros((K, M) if trans else (M, K), device="cuda")
        B = torch.zeros((K, N), device="cuda")
        # A[m,k] = 1 + m + 1000*k ; B = identity-ish: picks column k
        mm = torch.arange(M, device="cuda").float(); kk = torch.arange(K, device="cuda").float()
        Amk = (1 + mm[:, None] + 1000 * kk[None, :])
        A.copy_(Amk.t() if trans else Amk)
        for k in range(min(K, N)): B[k, k] = 1.0
    dA, dB = pkg.to_colmajor(A), pkg.to_colmajor(B)
    D = pkg.colmajor(M, N); D.fill_(-7.0)
    ctx.gemm_tf32x3(dA, dB, D, trans_a=trans); ctx.synchronize()
    ref = (A.double().t() if trans else A.double()) @ B.double()
    err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"trans={trans} M={M} N={N} K={K} {pattern}: rel max err {err:.3e}")
    if err > 1e-5:
        torch.set_printoptions(linewidth=200, precision=4, sci_mode=False)
        print(" D[:6,:6]=\n", D[:6, :6].cpu()); print(" ref[:6,:6]=\n", ref[:6, :6].float().cpu())
        print(" D[32:36,:4]=\n", D[32:36, :4].cpu()); print(" ref[32:36,:4]=\n", ref[32:36, :4].float().cpu())
        print(" nonzero frac", (D != 0).float().mean().item(), " untouched frac", (D == -7.0).float().mean().item())
    return err

for pat in ["idx", "rand"]:
    for (t, M, N, K) in [(True, 128, 128, 32), (True, 128, 128, 8), (False, 128, 128, 32), (False, 128, 128, 8),
                          (True, 128, 256, 64), (False, 256, 256, 64), (True, 256, 512, 4096), (False, 4096, 512, 256),
                          (True, 100, 70, 1000), (False, 1000, 70, 100)]:
        try:
            run(t, M, N, K, pat)
        except Exception as e:
            print("EXC", t, M, N, K, e)
