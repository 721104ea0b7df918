mkdir -p gpurun_out/r02
timeout 120 python - <<'PY'
import importlib, sys, torch
sys.path.insert(0, '.')
pkg = importlib.import_module("cuda-qr_b200")
ctx = pkg.Context(0); ctx.use_torch_stream()
for m, n in [(4096, 2048), (8192, 8192), (12000, 3000)]:
    A0 = torch.rand((n, m), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    A = pkg.colmajor(m, n); A.copy_(A0.t()); tau = torch.zeros(n, device="cuda")
    ctx.geqrf(A, tau); ctx.synchronize()
    Rd = torch.triu(A[:n].double()); G = A0.double() @ A0.double().t()
    print(m, n, "gram", float((Rd.t() @ Rd - G).norm() / G.norm()), flush=True)
PY
run() { echo "== bench $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'res', d['residual'], d['roofline']['by_class_ms'])"; }
run CQR_X=0
run CQR_SLICE_FUSED_ROWS=0
run CQR_SLICE_FUSED_ROWS=8192
for r in 8192 10240 12288 16384; do run CQR_PWS_ROWS=$r; done
timeout 1500 python -m pytest tests -m gpu -x -q -k "geqrf or square or legacy or partial or pair or form_q or apply_q or solve or chunked" > gpurun_out/r02/gputests_fs.log 2>&1
tail -4 gpurun_out/r02/gputests_fs.log
