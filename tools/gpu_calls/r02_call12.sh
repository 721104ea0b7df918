set -x
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -x -q -k "f64" > gpurun_out/r02/gputests_f64.log 2>&1
tail -15 gpurun_out/r02/gputests_f64.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2300 --csv --log-file gpurun_out/r02/launches_geqrf16384.csv python tools/one_geqrf.py 16384 > gpurun_out/r02/ncu_one.log 2>&1
python tools/launch_summary.py gpurun_out/r02/launches_geqrf16384.csv > gpurun_out/r02/launches_geqrf16384.txt; head -24 gpurun_out/r02/launches_geqrf16384.txt
python - <<'PY'
import importlib, sys, time, torch
sys.path.insert(0, '.')
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
for m, n in [(4096, 4096), (8192, 8192), (16384, 4096)]:
    A0 = torch.rand((n, m), device="cuda", dtype=torch.float64)
    A = pkg.colmajor(m, n, dtype=torch.float64); tau = torch.zeros(n, device="cuda", dtype=torch.float64)
    for _ in range(2):
        A.copy_(A0.t()); ctx.dgeqrf(A, tau)
    torch.cuda.synchronize()
    A.copy_(A0.t())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.dgeqrf(A, tau); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fl = 2.0 * m * n * n - 2.0 * n ** 3 / 3
    print(f"dgeqrf {m}x{n}: {ms:.1f} ms  {fl / ms / 1e9:.2f} TFLOP/s fp64")
PY
