"""Chain-stream view of one legacy mmqr(host buffers) call at 16384^2: CQR_LEGACY_TIMELINE dumps the library's event brackets;
this prints, per group of 32 panels, the panel time and the gap between consecutive panels (inner updates + waiting).
   python tools/e2e_timeline.py [out.txt]      (CQR_H2D_OVERLAP=0 for the blocking upload)"""
import importlib, os, sys, time
import numpy as np, torch
path = sys.argv[1] if len(sys.argv) > 1 else "/tmp/cqr_timeline.txt"
os.environ["CQR_LEGACY_TIMELINE"] = path
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
n = 16384
host = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
src = torch.rand((n, n), generator=torch.Generator().manual_seed(1))
host.copy_(src); hnp = host.numpy().T
tau = np.empty(pkg.tau_size(n, n), dtype=np.float32)
pkg.mmqr(hnp, tau)
for rep in range(2):
    host.copy_(src)
    t0 = time.perf_counter(); pkg.mmqr(hnp, tau); dt = time.perf_counter() - t0
rows = np.loadtxt(path)
names = ("panel", "gemm_tn", "gemm_nn", "misc", "chain_tn", "chain_nn", "chain_misc")
print(f"mmqr wall {dt * 1e3:.2f} ms; brackets {len(rows)}; span {rows[:, 1].max():.2f} ms")
for k, nm in enumerate(names):
    sel = rows[rows[:, 2] == k]
    print(f"  {nm:10s} n={len(sel):5d} busy {np.sum(sel[:, 1] - sel[:, 0]):8.2f} ms  first {sel[:, 0].min() if len(sel) else 0:7.2f}  last {sel[:, 1].max() if len(sel) else 0:7.2f}")
pan = rows[rows[:, 2] == 0]
pan = pan[np.argsort(pan[:, 0])]
for g in range(0, len(pan), 32):
    blk = pan[g:g + 32]
    gaps = blk[1:, 0] - blk[:-1, 1]
    print(f"  panels {g:3d}..{g + len(blk) - 1:3d}: start {blk[0, 0]:7.2f} ms  mean panel {np.mean(blk[:, 1] - blk[:, 0]) * 1e3:6.1f} us  mean gap {np.mean(gaps) * 1e3:6.1f} us  max gap {np.max(gaps) * 1e3:7.1f} us")
