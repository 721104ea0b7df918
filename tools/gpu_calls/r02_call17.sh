mkdir -p gpurun_out/r02
run() { echo "== $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --no-e2e --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), d['roofline']['by_class_ms'])"; }
{
run CQR_X=0
run CQR_PWS_ROWS=16384
run CQR_PWS_ROWS=12288
run CQR_TCHAIN_ROWS=8192
run CQR_TCHAIN_ROWS=16384
run CQR_PGROUPS_SMALL=1
run CQR_PGROUPS_SMALL=1 CQR_PWS_ROWS=16384
run CQR_PWS_ROWS=16384 CQR_TCHAIN_ROWS=16384
run CQR_PANEL_PAIR_MIN_ROWS=1024
run CQR_PARTIAL_OVERLAP_COLS=4096
} > gpurun_out/r02/knob_sweep_device.txt 2>&1
cat gpurun_out/r02/knob_sweep_device.txt
rune() { echo "== $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
{
rune CQR_CATCH_COLS=1024
rune CQR_CATCH_COLS=512
rune CQR_CATCH_COLS=1024 CQR_CATCH_CTAS_PCT=60
rune CQR_CATCH_COLS=4096 CQR_CATCH_CTAS_PCT=70
} > gpurun_out/r02/e2e_sweep3.txt 2>&1
cat gpurun_out/r02/e2e_sweep3.txt
