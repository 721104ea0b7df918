mkdir -p gpurun_out/r02
run() { echo "== bench $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'res', d['residual'], d['roofline']['by_class_ms'])"; }
run CQR_X=0
run CQR_PWS_INTERLEAVE=0
run CQR_X=0
run CQR_PWS_INTERLEAVE=0
echo "== timeline"; timeout 120 python tools/timeline.py 16384 28.0 30.2 > gpurun_out/r02/timeline_mid.txt 2>&1; head -75 gpurun_out/r02/timeline_mid.txt
timeout 1500 python -m pytest tests -m gpu -x -q -k "geqrf or square or legacy or partial or pair or form_q or apply_q or solve or chunked" > gpurun_out/r02/gputests_il.log 2>&1
tail -4 gpurun_out/r02/gputests_il.log
