run() { echo "== bench $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --no-e2e --steps 6 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2))"; }
for r in 12288 10240 8192 6144 14336 12288; do run CQR_TCHAIN_ROWS=$r; done
