set -x
mkdir -p gpurun_out/r02
timeout 300 python tools/check_tsqr_mma.py > gpurun_out/r02/check_tsqr_mma.txt 2>&1
cat gpurun_out/r02/check_tsqr_mma.txt
timeout 300 python tools/tsqr_bench.py > gpurun_out/r02/tsqr_bench_mma.txt 2>&1
cat gpurun_out/r02/tsqr_bench_mma.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "tsqr" > gpurun_out/r02/gputests_tsqr.log 2>&1
tail -5 gpurun_out/r02/gputests_tsqr.log
