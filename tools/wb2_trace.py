"""Debug aid: phase timing of panel_wb2_kernel (two pivot columns per exchange) from a trace build
(make -C cuda-qr_b200/csrc clean; make -C cuda-qr_b200/csrc TRACE=1).   python tools/wb2_trace.py [m ...]
Per exchange step, lane 0 of every warp of CTA 0 and of the last CTA records clock64 at: 0 step entry, 1 x/y published and
loaded, 2 dots + shuffles done, 3 past the CTA barrier, 4 (warp 0) phase-1 mbarrier passed, 5 totals in (phase 2 / two-
cluster flags), 6 reflector scalars done, 7 update + column stores done.  Kernel marks: entry, panel loaded, steps done,
results stored, T done."""
import ctypes, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
if not hasattr(pkg.lib, "cqr_debug_wb2_trace"):
    sys.exit("libcudaqr_b200.so was not built with TRACE=1")
ctx = pkg.Context(0); ctx.use_torch_stream()
names = ["entry -> x,y published+loaded", "dots + shuffles", "CTA barrier", "phase 1 (warp 0)", "phase 2 / totals in",
         "reflector scalars", "update + stores"]
for m in [int(a) for a in sys.argv[1:]] or [2048, 8192]:
    A0 = pkg.colmajor(m, 64); A0.copy_(torch.rand((m, 64), device="cuda"))
    A = pkg.colmajor(m, 64); tau = torch.zeros(64, device="cuda")
    for _ in range(3):
        A.copy_(A0); ctx.geqrf(A, tau)
    ctx.synchronize()
    steps = np.zeros((2, 8, 64, 8), dtype=np.int64); marks = np.zeros((2, 8, 5), dtype=np.int64)
    pkg.lib.cqr_debug_wb2_trace(steps.ctypes.data_as(ctypes.c_void_p), marks.ctypes.data_as(ctypes.c_void_p))
    nex = int((steps[0, 0, :, 0] > 0).sum())            # exchanges of the last launch (32 without fallbacks)
    for ci, cname in enumerate(["CTA 0", "last CTA"]):
        t = steps[ci][:, :nex]                           # [warp][exchange][8]
        live = [w for w in range(8) if t[w, 0, 0] > 0]
        if not live:
            continue
        t = t[live]
        per = (t[:, 1:, 0] - t[:, :-1, 0]).mean()
        print(f"m={m} {cname}: {nex} exchanges, {len(live)} warps, mean cycles per exchange step {per:.0f}")
        for k in range(7):
            if k == 3:                                   # phase 1 is recorded by warp 0 only
                d = t[:1, :, 4] - t[:1, :, 3]
            elif k == 4:
                d = t[:, :, 5] - np.where(t[:, :, 4] > t[:, :, 3], t[:, :, 4], t[:, :, 3])
            else:
                d = t[:, :, k + 1] - t[:, :, k] if k < 3 else t[:, :, k + 1] - t[:, :, k]
            print(f"    {names[k]:32s} mean {d.mean():7.0f}  min {d.min():6d}  max {d.max():6d}")
        mk = marks[ci][live]
        lab = ["load panel", "all steps", "store A and V", "T (CTA 0)"]
        print("    kernel: " + ",  ".join(f"{lab[k]} {int((mk[:, k + 1] - mk[:, k]).mean())}" for k in range(4)) + "  cycles")
