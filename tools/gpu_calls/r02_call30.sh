mkdir -p gpurun_out/r02
export CQR_PANEL_BENCH_MODES=1
echo "== digest old (trace build of the previous kernel)"; CQR_LIB=cuda-qr_b200/csrc/build/trace/libcudaqr_b200.so timeout 120 python tools/panel_digest.py > gpurun_out/r02/digest_old.txt 2>&1; tail -3 gpurun_out/r02/digest_old.txt
echo "== digest new"; timeout 120 python tools/panel_digest.py > gpurun_out/r02/digest_new.txt 2>&1; tail -3 gpurun_out/r02/digest_new.txt
diff gpurun_out/r02/digest_old.txt gpurun_out/r02/digest_new.txt && echo "DIGESTS IDENTICAL"
echo "== panel bench new"; timeout 120 python tools/panel_bench.py 512 2048 4096 8192 16384
timeout 200 python bench.py --no-extra --no-cpu --steps 4 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'residual', d['residual'], d['roofline']['by_class_ms'])"
timeout 1500 python -m pytest tests -m gpu -x -q -k "geqrf or square or legacy or partial or pair or form_q or apply_q or solve" > gpurun_out/r02/gputests_rot.log 2>&1
tail -4 gpurun_out/r02/gputests_rot.log
