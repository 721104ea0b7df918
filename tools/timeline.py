"""Per-launch-group timeline of one geqrf (library event brackets): prints the brackets of a few outer blocks so the
serial path (panel chain -> block T -> look-ahead slice) can be read off.   python tools/timeline.py [n] [first_ms] [last_ms]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
lo = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
hi = float(sys.argv[3]) if len(sys.argv) > 3 else 23.5
A0 = pkg.colmajor(n, n); A0.copy_(torch.rand((n, n), device="cuda"))
A = pkg.colmajor(n, n); tau = torch.zeros(n, device="cuda")
for _ in range(2):
    A.copy_(A0); ctx.geqrf(A, tau)
torch.cuda.synchronize()
A.copy_(A0); torch.cuda.synchronize()
ctx.profile_begin(); ctx.geqrf(A, tau)
tl = ctx.profile_timeline()
prof = ctx.profile_end()
print(f"{len(tl)} brackets, last end {max(t[1] for t in tl):.2f} ms")
chain = [t for t in tl if t[2] in ("panel", "chain_tn", "chain_nn", "chain_misc")]
main = [t for t in tl if t not in chain]
def busy(xs):
    return sum(t[1] - t[0] for t in xs)
print(f"chain stream busy {busy(chain):.2f} ms, GEMM stream busy {busy(main):.2f} ms")
# per outer block (four panels): period, time in panel kernels, in chain-class brackets (inner updates and look-ahead slices,
# on either stream) and in the GEMM stream's own brackets -- shows which of the two streams is the busy one where
panels = sorted(t for t in tl if t[2] == "panel")
print("block  first_panel_ms  period_us  panels_us  chain_us  gemm_us")
for b in range(0, len(panels) - 4, 4):
    w0, w1 = panels[b][0], panels[b + 4][0]
    inw = [t for t in tl if w0 <= t[0] < w1]
    pan = sum(t[1] - t[0] for t in inw if t[2] == "panel")
    ch = sum(t[1] - t[0] for t in inw if t[2] in ("chain_tn", "chain_nn", "chain_misc"))
    gm = sum(t[1] - t[0] for t in inw if t[2] not in ("panel", "chain_tn", "chain_nn", "chain_misc"))
    if (b // 4) % 4 == 0:
        print(f"{b // 4:5d}  {w0:14.2f}  {(w1 - w0) * 1e3:9.0f}  {pan * 1e3:9.0f}  {ch * 1e3:8.0f}  {gm * 1e3:7.0f}")
for t0, t1, c in sorted(tl):
    if lo <= t0 <= hi:
        lane = "P" if c in ("panel", "chain_tn", "chain_nn", "chain_misc") else "G"
        print(f"{lane} {t0:9.3f} -> {t1:9.3f}  ({(t1 - t0) * 1e3:7.1f} us)  {c}")
