export CQR_PANEL_BENCH_MODES=1
run() { echo "== bench $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['roofline']['by_class_ms']['panel'])"; }
for rep in 1 2; do
echo "== rotated (default lib)"; timeout 120 python tools/panel_bench.py 2048 4096 8192 16384
run CQR_X=0
echo "== 8 instantiations (norot lib)"; CQR_LIB=cuda-qr_b200/csrc/build/norot/libcudaqr_b200.so timeout 120 python tools/panel_bench.py 2048 4096 8192 16384
run CQR_LIB=cuda-qr_b200/csrc/build/norot/libcudaqr_b200.so
done
