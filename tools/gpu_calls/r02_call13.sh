set -x
mkdir -p gpurun_out/r02
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02/gputests_full.log 2>&1
tail -6 gpurun_out/r02/gputests_full.log
timeout 900 python bench.py > gpurun_out/r02/bench_n1.json 2> gpurun_out/r02/bench_n1.err
tail -c 300 gpurun_out/r02/bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r02/bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['tsqr']['ms_per_step'], d['batched']['ms_per_step'], d['residual'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python - <<'PY'
import importlib, sys, torch
sys.path.insert(0, '.')
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
for m, n in [(4096, 4096), (16384, 4096)]:
    A0 = torch.rand((n, m), device="cuda", dtype=torch.float64)
    A = pkg.colmajor(m, n, dtype=torch.float64); tau = torch.zeros(n, device="cuda", dtype=torch.float64)
    for _ in range(2):
        A.copy_(A0.t()); ctx.dgeqrf(A, tau)
    torch.cuda.synchronize(); A.copy_(A0.t())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.dgeqrf(A, tau); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); fl = 2.0 * m * n * n - 2.0 * n ** 3 / 3
    print(f"dgeqrf {m}x{n}: {ms:.1f} ms  {fl / ms / 1e9:.2f} TFLOP/s fp64")
PY
