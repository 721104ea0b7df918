mkdir -p gpurun_out/r02
run() { echo "== $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
{
run CQR_X=0
run CQR_H2D_JOIN=model
run CQR_CATCH_CTAS_PCT=50
run CQR_CATCH_CTAS_PCT=30
} > gpurun_out/r02/e2e_sweep2.txt 2>&1
cat gpurun_out/r02/e2e_sweep2.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "chunked_upload or tensor_pipe_leaf" 2>&1 | tail -3
