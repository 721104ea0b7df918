"""The reference's own benchmark grid (timing.txt: seconds per mmqr call on host buffers, transfers included, mean of 3
trials, sizes rounded by qr.cu:722-734) re-run through this library's legacy `mmqr` entry point, next to the
reference's published MMQR and MAGMA columns (Kepler-class GPU, MAGMA 2.0.2).   python tools/timing_table.py > table.md"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
import oracle

REF = {  # timing.txt: requested (m, n) -> (MMQR s, MAGMA s)
    (256, 64): (0.017642, 0.022984), (512, 64): (0.034541, 0.023280), (1024, 64): (0.068002, 0.024406),
    (2048, 64): (0.135096, 0.025067), (4096, 64): (0.269188, 0.028084), (8192, 64): (0.545109, 0.033503),
    (16384, 64): (1.094346, 0.044161), (32768, 64): (2.189796, 0.066345), (65536, 64): (4.396491, 0.113676),
    (131072, 64): (8.793325, 0.249329),
    (64, 64): (0.006715, 0.063225), (128, 128): (0.021271, 0.023507), (256, 256): (0.073523, 0.028029),
    (512, 512): (0.268534, 0.029531), (1024, 1024): (1.168431, 0.044149), (2048, 2048): (4.656755, 0.097721),
    (4096, 4096): (24.307268, 0.305895),
}


def round_like_reference(m, n, PR=64, PC=4):
    m = PR + int((m - PR) / (PR - PC) + 0.5) * (PR - PC)
    k = int(n / PC + 0.5) or 1
    n = k * PC
    while n > m:
        n -= PC
    return m, n


print("| requested | exact (qr.cu:722-734) | reference MMQR s | reference MAGMA s | this library s | GFLOP/s | speed-up vs MMQR | vs MAGMA |")
print("|---|---|---|---|---|---|---|---|")
pkg.mmqr(np.asfortranarray(np.random.rand(128, 64).astype(np.float32)))          # context + first-touch costs out of the table
for (mr, nr), (t_mmqr, t_magma) in REF.items():
    m, n = round_like_reference(mr, nr)
    A = oracle.rand_matrix(m, n, 12) if m * n <= (1 << 22) else np.asfortranarray(np.random.default_rng(12).random((m, n), dtype=np.float32))
    RV = A.copy(order="F")
    tau = np.empty(pkg.tau_size(m, n), dtype=np.float32)
    pkg.mmqr(RV, tau)                                                               # warm-up at this size (workspace growth)
    ts = []
    for _ in range(3):
        RV[...] = A
        t0 = time.perf_counter(); pkg.mmqr(RV, tau); ts.append(time.perf_counter() - t0)
    t = sum(ts) / 3
    gf = (2.0 * m * n * n - 2.0 * n ** 3 / 3) / t / 1e9
    print(f"| {mr}x{nr} | {m}x{n} | {t_mmqr:.6f} | {t_magma:.6f} | {t:.6f} | {gf:.1f} | {t_mmqr / t:.0f}x | {t_magma / t:.1f}x |", flush=True)
