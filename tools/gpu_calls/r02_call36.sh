mkdir -p gpurun_out/r02
timeout 120 python tools/timeline.py 16384 28.0 29.3 > gpurun_out/r02/timeline_after.txt 2>&1; head -60 gpurun_out/r02/timeline_after.txt
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02/gputests_full.log 2>&1; tail -5 gpurun_out/r02/gputests_full.log
timeout 600 python bench.py > gpurun_out/r02/bench_n1.json 2> gpurun_out/r02/bench_n1.err; tail -c 1500 gpurun_out/r02/bench_n1.json
