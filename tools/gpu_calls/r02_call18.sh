mkdir -p gpurun_out/r02
rune() { echo "== $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
{
rune CQR_X=0
rune CQR_H2D_SLICE_ON_CHAIN=0
rune CQR_X=1
} > gpurun_out/r02/e2e_sweep4.txt 2>&1
cat gpurun_out/r02/e2e_sweep4.txt
timeout 300 python -m pytest tests -m gpu -x -q -k "chunked_upload" 2>&1 | tail -2
timeout 300 python tools/e2e_timeline.py /tmp/tl_overlap.txt 2>&1 | tail -9
