mkdir -p gpurun_out
CQR_FLAT_CFG=2 python -m pytest tests -m gpu -x -q -k "tsqr" 2>&1 | tail -5
for c in 0 1 2; do CQR_FLAT_CFG=$c python tools/tsqr_bench.py 8388608 1048576 2>&1 | grep "flat=1"; done
