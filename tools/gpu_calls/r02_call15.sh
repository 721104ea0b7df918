mkdir -p gpurun_out/r02
CHECK_MODE=3 timeout 300 python tools/check_tsqr_mma.py > gpurun_out/r02/check_tsqr_pair.txt 2>&1
cat gpurun_out/r02/check_tsqr_pair.txt
timeout 300 python tools/tsqr_bench.py 8388608 1048576 > gpurun_out/r02/tsqr_bench_pair.txt 2>&1
cat gpurun_out/r02/tsqr_bench_pair.txt
run() { echo "== $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
run CQR_X=0
run CQR_CATCH_COLS=2048
timeout 300 python -m pytest tests -m gpu -x -q -k "chunked_upload" 2>&1 | tail -2
