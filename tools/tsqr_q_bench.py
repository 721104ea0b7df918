"""TSQR with implicit Q (cqr_tsqr_factor) and thin-Q expansion (cqr_tsqr_form_q): flat leaf vs 256-row tile leaves."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8388608
A0 = pkg.colmajor(m, 64); A0.copy_(torch.rand((m, 64), device="cuda", generator=torch.Generator(device="cuda").manual_seed(12)))
A = pkg.colmajor(m, 64); R = pkg.colmajor(64, 64); Q = pkg.colmajor(m, 64)
def timed(fn, reps=8):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
for flat in (1, 0):
    ctx.set_option(pkg.OPT_FLAT_TSQR, flat)
    A.copy_(A0); ctx.tsqr_factor(A, R); ctx.tsqr_form_q(Q); ctx.synchronize()
    tf = []
    for _ in range(6):
        A.copy_(A0); tf.append(timed(lambda: ctx.tsqr_factor(A, R), 1))
    tf.sort()
    tq = timed(lambda: ctx.tsqr_form_q(Q))
    orth = float((Q.t().double() @ Q.double() - torch.eye(64, device="cuda", dtype=torch.float64)).norm()) / (64 * 2.0 ** -23)
    be = float((A0.double() - Q.double() @ torch.triu(R.double())).norm() / A0.double().norm()) / (64 * 2.0 ** -23)
    print(f"tsqr {m} x 64 flat={flat}: factor (implicit Q) {tf[len(tf)//2]:.3f} ms, form thin Q {tq:.3f} ms, orth {orth:.3f} backward {be:.3f} (units of n eps)", flush=True)
