"""Where a warp's time goes in the tensor-pipe TSQR leaf (trace build: make -C cuda-qr_b200/csrc clean && make MMATRACE=1):
clock64 brackets around the sub-panel factorisations and the trailing products of warp 0 of CTA 0.
   python tools/mma_trace.py [rows]"""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8388608
A = pkg.colmajor(m, 64); A.copy_(torch.rand((m, 64), device="cuda"))
R = pkg.colmajor(64, 64)
buf = (ctypes.c_longlong * 4)()
ctx.tsqr_r(A, R); ctx.synchronize()
pkg.lib.cqr_debug_mma_trace(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctx.tsqr_r(A, R); e1.record(); ctx.synchronize()
pkg.lib.cqr_debug_mma_trace(buf)
blocks = max(1, buf[3])
print(f"{m} x 64: {e0.elapsed_time(e1):.3f} ms; warp 0 of CTA 0 (all levels): {blocks} block steps, per block: "
      f"sub-panels {buf[1] / blocks:.0f} clk, trailing {buf[2] / blocks:.0f} clk")
