set -x
mkdir -p gpurun_out/r02
(python bench.py --config1-only --config1-full > gpurun_out/r02/config1_full.json 2> gpurun_out/r02/config1_full.err) &
CFG1=$!
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02/gputests_call2.log 2>&1
tail -15 gpurun_out/r02/gputests_call2.log
timeout 600 python bench.py > gpurun_out/r02/bench_call2.json 2> gpurun_out/r02/bench_call2.err
tail -c 600 gpurun_out/r02/bench_call2.err
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02/racecheck_small_after_fix.log 2>&1
tail -3 gpurun_out/r02/racecheck_small_after_fix.log
wait $CFG1
cat gpurun_out/r02/config1_full.json
