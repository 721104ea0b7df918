run() { echo "== bench $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 5 --warmup 2 --e2e-steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'res', d['residual'])"; }
run CQR_SLICE_CHAIN_ROWS=0
run CQR_X=0
run CQR_SLICE_CHAIN_ROWS=14336
run CQR_SLICE_CHAIN_ROWS=10240 CQR_TCHAIN_ROWS=10240
run CQR_SLICE_CHAIN_ROWS=0
run CQR_X=0
