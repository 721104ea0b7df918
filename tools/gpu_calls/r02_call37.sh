run() { echo "== bench $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
run CQR_X=0
for g in 3 4; do for r in 4096 6144 8192; do run CQR_PGROUPS_SMALL=$g CQR_CHAIN_FUSED_ROWS=$r; done; done
run CQR_PGROUPS_BIG=3
run CQR_PGROUPS_BIG=3 CQR_PGROUPS_SMALL=3 CQR_CHAIN_FUSED_ROWS=6144
