run() { echo "== bench $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 2 --warmup 2 --e2e-steps 4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'res', d['residual'])"; }
for t in 0 1 2 4 8 0 2; do run CQR_CATCH_TILES=$t; done
