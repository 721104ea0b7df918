run() { echo "== bench $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 2 --warmup 2 --e2e-steps 4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
for t in 2 3; do for cc in 2048 4096; do run CQR_CATCH_TILES=$t CQR_CATCH_COLS=$cc; done; done
run CQR_CATCH_TILES=2 CQR_H2D_JOIN=model
run CQR_CATCH_TILES=2 CQR_CATCH_CTAS_PCT=50
