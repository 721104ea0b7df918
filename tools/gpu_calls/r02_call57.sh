mkdir -p gpurun_out/r02
export CQR_PANEL_BENCH_MODES=1
timeout 120 python tools/panel_digest.py > gpurun_out/r02/digest_new.txt 2>&1; diff tools/gpu_calls/digest_ref.txt gpurun_out/r02/digest_new.txt && echo "DIGESTS IDENTICAL"
timeout 120 python tools/panel_bench.py 8192 10240 12288 16384
run() { echo "== bench $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'res', d['residual'], d['roofline']['by_class_ms']['panel'])"; }
run CQR_X=0
timeout 900 python -m pytest tests -m gpu -x -q -k "geqrf or square or legacy or partial or pair or chunked" > gpurun_out/r02/gputests_x2.log 2>&1; tail -3 gpurun_out/r02/gputests_x2.log
