set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "tsqr" 2>&1 | tail -15
CQR_FLAT_MINB=3 python tools/tsqr_bench.py 2>&1 | tail -8
CQR_FLAT_MINB=2 python tools/tsqr_bench.py 8388608 1048576 2>&1 | tail -4
ncu --set full --import-source on --clock-control none -k regex:tsqr_flat -c 1 -o gpurun_out/r01_tsqr_flat python tools/tsqr_bench.py once 8388608 > gpurun_out/ncu_flat.log 2>&1
tail -3 gpurun_out/ncu_flat.log
