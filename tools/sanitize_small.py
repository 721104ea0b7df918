"""Small-shape workload for compute-sanitizer (memcheck / racecheck) over the warp-resident kernels and the legacy path:
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
g = torch.Generator(device="cuda").manual_seed(1)
for m, n in [(16444, 64), (20011, 17)]:                    # flat TSQR leaf: aligned and ragged / unaligned
    A = pkg.colmajor(m, n); A.copy_(torch.rand((m, n), device="cuda", generator=g))
    R = pkg.colmajor(n, n); ctx.tsqr_r(A, R); ctx.synchronize()
    G = A.t().double() @ A.double(); Rd = torch.triu(R.double())
    print("tsqr_r", m, n, float((Rd.t() @ Rd - G).norm() / G.norm()))
# Gram leaf (gram_umma.cu: TMA ring, converter / MMA / epilogue warps, finish kernel) with several groups per CTA, a narrow
# matrix, and a singular input that raises the gate so the Householder leaf behind it runs as well
for m, n, dup in [(148 * 128 * 3 + 256, 64, False), (16384, 24, False), (32768, 64, True)]:
    A = pkg.colmajor(m, n); A.copy_(torch.rand((m, n), device="cuda", generator=g))
    if dup:
        A[:, 5] = A[:, 2]
    R = pkg.colmajor(n, n); ctx.tsqr_r(A, R); ctx.synchronize()
    G = A.t().double() @ A.double(); Rd = torch.triu(R.double())
    print("tsqr_r gram leaf", m, n, ctx.tsqr_gram_info(), float((Rd.t() @ Rd - G).norm() / G.norm()))
for m, n, batch in [(64, 64, 40), (50, 33, 9)]:            # batched warp kernel
    A3 = torch.rand((batch, n, m), device="cuda", generator=g); tau = torch.zeros((batch, n), device="cuda")
    O = A3.clone(); ctx.geqrf_batched(A3, tau); ctx.synchronize()
    Rd = torch.triu(A3.transpose(1, 2)[:, :n, :].double()); G = O.double() @ O.double().transpose(1, 2)
    print("batched", m, n, float(((Rd.transpose(1, 2) @ Rd - G).flatten(1).norm(dim=1) / G.flatten(1).norm(dim=1)).max()))
A = np.asfortranarray(np.random.default_rng(0).random((300, 200), dtype=np.float32))
RV = A.copy(order="F"); tau = pkg.mmqr(RV); Q, R = pkg.explicitQR(RV, tau)
print("legacy 300x200 residual", float(np.linalg.norm(Q.astype(np.float64) @ R.astype(np.float64) - A) / np.linalg.norm(A)))
# panel kernels on the cluster exchange (panel_wb2.cu from 2048 rows up); the wider cases run the look-ahead schedule with
# inner updates -- under CQR_CHAIN_FUSED=2 through the one-launch K = 64 update (chain_update.cu)
for m, n in [(2500, 64), (4100, 128), (3000, 600)]:
    A = pkg.colmajor(m, n); A.copy_(torch.rand((m, n), device="cuda", generator=g)); O = A.clone()
    tau = torch.zeros(n, device="cuda"); ctx.geqrf(A, tau); ctx.synchronize()
    Rd = torch.triu(A[:n].double()); G = O.t().double() @ O.double()
    print("geqrf", m, n, float((Rd.t() @ Rd - G).norm() / G.norm()))
