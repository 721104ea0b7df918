export CQR_PANEL_BENCH_MODES=1
R="512 2048 4096 8192 16384"
run() { echo "== bench $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 4 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'residual', d['residual'], d['roofline']['by_class_ms'])"; }
echo "== blocked T (default lib)"; timeout 120 python tools/panel_bench.py $R
run CQR_X=0
echo "== blocked T + scalar FMA"; CQR_LIB=cuda-qr_b200/csrc/build/scalar/libcudaqr_b200.so timeout 120 python tools/panel_bench.py $R
run CQR_LIB=cuda-qr_b200/csrc/build/scalar/libcudaqr_b200.so
timeout 1500 python -m pytest tests -m gpu -x -q -k "geqrf or square or legacy or partial or pair or form_q or apply_q or solve" > gpurun_out/r02/gputests_tblock.log 2>&1
tail -4 gpurun_out/r02/gputests_tblock.log
