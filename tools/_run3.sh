mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
python tools/tsqr_bench.py 8388608 1048576 2>&1 | grep "flat=1"
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_mid.json 2> gpurun_out/bench_r01_mid.err; tail -c 3000 gpurun_out/bench_r01_mid.json
