mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "batched" 2>&1 | tail -15
python tools/batched_bench.py
CQR_BATCHED_CTA=1 python tools/batched_bench.py
