"""Debug aid: phase timing of panel_hh_kernel from a -DCQR_HH_TRACE build (make -C cuda-qr_b200/csrc TRACE=1).
    python tools/hh_trace.py m"""
import ctypes, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
for m in [int(a) for a in sys.argv[1:]] or [64, 512, 16384]:
    A0 = pkg.colmajor(m, 64); A0.copy_(torch.rand((m, 64), device="cuda"))
    A = pkg.colmajor(m, 64); tau = torch.zeros(64, device="cuda")
    for _ in range(3):
        A.copy_(A0); ctx.geqrf(A, tau)
    ctx.synchronize()
    buf = np.zeros((2, 16, 64, 6), dtype=np.int64)
    pkg.lib.cqr_debug_hh_trace(buf.ctypes.data_as(ctypes.c_void_p))
    names = ["sync1->dots", "dots->publish+gather", "gather->(owner scalars)", "wait sync2", "update", "next sync1"]
    for ci, cname in enumerate(["CTA 0", "last CTA"]):
        t = buf[ci]                      # [warp][step][6]
        step = (t[:, 1:, 0] - t[:, :-1, 0]).mean()
        print(f"m={m} {cname}: mean cycles per step {step:.0f}")
        for k in range(5):
            d = (t[:, 1:-1, k + 1] - t[:, 1:-1, k])
            print(f"    {names[k]:28s} mean {d.mean():7.0f}  min {d.min():6d}  max {d.max():6d}")
        d = t[:, 2:, 0] - t[:, 1:-1, 5]
        print(f"    {'update end -> next sync1':28s} mean {d.mean():7.0f}  min {d.min():6d}  max {d.max():6d}")
