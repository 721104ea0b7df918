run() { echo "== bench $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'res', d['residual'], d['gpu_launches'])"; }
run CQR_CHAIN_COOP=1
run CQR_CHAIN_COOP=0
run CQR_CHAIN_COOP=1
run CQR_CHAIN_COOP=0
