mkdir -p gpurun_out/r02
export CQR_PANEL_BENCH_MODES=1
timeout 120 python tools/panel_digest.py > gpurun_out/r02/digest_new.txt 2>&1; diff tools/gpu_calls/digest_ref.txt gpurun_out/r02/digest_new.txt && echo "DIGESTS IDENTICAL"
run() { echo "== bench $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['roofline']['by_class_ms']['panel'])"; }
echo "== rotated + two-pass update (default lib)"; timeout 120 python tools/panel_bench.py 2048 4096 8192 16384
run CQR_X=0
echo "== 8 instantiations (norot lib)"; CQR_LIB=cuda-qr_b200/csrc/build/norot/libcudaqr_b200.so timeout 120 python tools/panel_bench.py 2048 4096 8192 16384
run CQR_LIB=cuda-qr_b200/csrc/build/norot/libcudaqr_b200.so
echo "== timeline"; timeout 120 python tools/timeline.py 16384 30.0 32.2 > gpurun_out/r02/timeline_mid.txt 2>&1; head -70 gpurun_out/r02/timeline_mid.txt
