mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -x -q -k "peer" 2>&1 | tail -30
