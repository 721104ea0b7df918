"""Latency of one 64-column panel factorisation (cqr_geqrf on m x 64: zero-fill of V, panel kernel(s), T builder)
for both panel modes, CUDA events, 20 repetitions each.   python tools/panel_bench.py [m ...]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
ms_list = [int(a) for a in sys.argv[1:]] or [64, 512, 1024, 2048, 4096, 8192, 16384, 32768]
for m in ms_list:
    A0 = pkg.colmajor(m, 64); A0.copy_(torch.rand((m, 64), device="cuda"))
    A = pkg.colmajor(m, 64); tau = torch.zeros(64, device="cuda")
    out = []
    for mode in ((1,) if os.environ.get('CQR_PANEL_BENCH_MODES') == '1' else (1, 0)):
        ctx.set_option(pkg.OPT_PANEL, mode)
        for _ in range(3):
            A.copy_(A0); ctx.geqrf(A, tau)
        ctx.synchronize()
        ts = []
        for _ in range(20):
            A.copy_(A0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.geqrf(A, tau); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        assert ctx.last_error() == 0 if hasattr(ctx, 'last_error') else True
        out.append(f"mode {mode}: median {ts[10]:8.1f} us  min {ts[0]:8.1f} us")
    print(f"panel {m:6d} x 64   " + "   ".join(out), flush=True)
ctx.set_option(pkg.OPT_PANEL, 1)
