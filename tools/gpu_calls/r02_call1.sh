set -x
mkdir -p gpurun_out/r02
./tools/probes/mma_tf32_probe > gpurun_out/r02/mma_probe.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02/smi.txt
CQR_PANEL_PAIR_MAX_ROWS=16384 timeout 600 python -m pytest tests -m gpu -q -x -k "panel_hh or square_properties or pair_step" > gpurun_out/r02/pair16384_tests.log 2>&1
CQR_PANEL_PAIR_MAX_ROWS=16384 timeout 300 python tools/panel_bench.py 10240 12288 16384 > gpurun_out/r02/panel_bench_pair16384.txt 2>&1
timeout 300 python tools/panel_bench.py 2048 4096 8192 10240 12288 16384 > gpurun_out/r02/panel_bench_default.txt 2>&1
CQR_PANEL_PAIR_MAX_ROWS=16384 timeout 300 python bench.py --no-extra --no-cpu --no-e2e --steps 3 > gpurun_out/r02/bench_pair16384.json 2> gpurun_out/r02/bench_pair16384.err
timeout 300 python bench.py --no-extra --no-cpu --no-e2e --steps 3 > gpurun_out/r02/bench_default.json 2> gpurun_out/r02/bench_default.err
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02/racecheck_small.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02/memcheck_small.log 2>&1
tail -3 gpurun_out/r02/*.log gpurun_out/r02/*.txt
