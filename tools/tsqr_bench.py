"""R-only TSQR (cqr_tsqr_r) timing: flat-tree warp leaf vs 256-row tile leaves, CUDA events, median of 15.
   python tools/tsqr_bench.py [rows ...]     (CQR_FLAT_MINB=2|3 picks the flat kernel's register budget)
   python tools/tsqr_bench.py once ROWS      (one flat call: the ncu target)"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
args = sys.argv[1:]
once = bool(args) and args[0] == "once"
if once:
    args = args[1:]
rows = [int(a) for a in args] or [8388608, 1048576, 131072]
for m in rows:
    A = pkg.colmajor(m, 64); A.copy_(torch.rand((m, 64), device="cuda", generator=torch.Generator(device="cuda").manual_seed(12)))
    R = pkg.colmajor(64, 64)
    if once:
        ctx.tsqr_r(A, R); ctx.synchronize()
        continue
    G = A.t().double() @ A.double()
    for flat in (4, 3, 2, 1, 0):       # 4 = Gram leaf (default), 3 = SIMT pair step, 2 = mma.sync leaf, 1 = SIMT flat leaf, 0 = tile leaves
        ctx.set_option(pkg.OPT_FLAT_TSQR, flat)
        for _ in range(3):
            ctx.tsqr_r(A, R)
        ctx.synchronize()
        ts = []
        l0 = ctx.launch_count()
        for _ in range(15):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.tsqr_r(A, R); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        launches = (ctx.launch_count() - l0) // 15
        ctx.profile_begin(); ctx.tsqr_r(A, R); prof = ctx.profile_end()
        ts.sort()
        Rd = torch.triu(R.double())
        gram = float((Rd.t() @ Rd - G).norm() / G.norm())
        flops = 2.0 * m * 64 * 64 - 2.0 * 64 ** 3 / 3
        print(f"tsqr_r {m:8d} x 64  flat={flat} minb={os.environ.get('CQR_FLAT_MINB', '3')}: median {ts[7]:7.3f} ms  min {ts[0]:7.3f} ms  "
              f"{flops / ts[7] / 1e9:7.1f} TFLOP/s  {4.0 * m * 64 / ts[7] / 1e6:7.1f} GB/s  launches {launches}  gram {gram:.2e}", flush=True)
    ctx.set_option(pkg.OPT_FLAT_TSQR, pkg.TSQR_LEAF_DEFAULT)
    del A
