run() { echo "== $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 4 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['gpu_launches'])"; }
for r in 0 2048 4096 6144 8192 10240; do run CQR_CHAIN_FUSED_ROWS=$r; done
