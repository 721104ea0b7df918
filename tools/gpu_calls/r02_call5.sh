set -x
mkdir -p gpurun_out/r02
timeout 300 python tools/check_tsqr_mma.py > gpurun_out/r02/check_tsqr_mma.txt 2>&1
cat gpurun_out/r02/check_tsqr_mma.txt
timeout 300 python tools/tsqr_bench.py 8388608 > gpurun_out/r02/tsqr_bench_mma.txt 2>&1
cat gpurun_out/r02/tsqr_bench_mma.txt
CQR_LIB=$PWD/cuda-qr_b200/libcudaqr_b200_trace.so timeout 120 python tools/mma_trace.py 8388608 > gpurun_out/r02/mma_trace.txt 2>&1
CQR_LIB=$PWD/cuda-qr_b200/libcudaqr_b200_trace.so timeout 120 python tools/mma_trace.py 65536 >> gpurun_out/r02/mma_trace.txt 2>&1
cat gpurun_out/r02/mma_trace.txt
