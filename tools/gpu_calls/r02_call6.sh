set -x
mkdir -p gpurun_out/r02
timeout 300 python tools/check_tsqr_mma.py > gpurun_out/r02/check_tsqr_mma.txt 2>&1
cat gpurun_out/r02/check_tsqr_mma.txt
timeout 300 python tools/tsqr_bench.py 8388608 > gpurun_out/r02/tsqr_bench_mma.txt 2>&1
cat gpurun_out/r02/tsqr_bench_mma.txt
CQR_LIB=$PWD/cuda-qr_b200/libcudaqr_b200_trace.so timeout 120 python tools/mma_trace.py 8388608 > gpurun_out/r02/mma_trace.txt 2>&1
CQR_LIB=$PWD/cuda-qr_b200/libcudaqr_b200_trace.so timeout 120 python tools/mma_trace.py 65536 >> gpurun_out/r02/mma_trace.txt 2>&1
cat gpurun_out/r02/mma_trace.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "chunked_upload or legacy" > gpurun_out/r02/gputests_legacy.log 2>&1
tail -5 gpurun_out/r02/gputests_legacy.log
timeout 300 python bench.py --no-extra --no-cpu --steps 3 > gpurun_out/r02/bench_e2e_overlap.json 2> gpurun_out/r02/bench_e2e_overlap.err
python -c "
import json; d=json.loads(open('gpurun_out/r02/bench_e2e_overlap.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'])"
CQR_H2D_OVERLAP=0 timeout 300 python bench.py --no-extra --no-cpu --steps 3 > gpurun_out/r02/bench_e2e_nooverlap.json 2> gpurun_out/r02/bench_e2e_nooverlap.err
python -c "
import json; d=json.loads(open('gpurun_out/r02/bench_e2e_nooverlap.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'])"
