mkdir -p gpurun_out/r02
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02/gputests_full.log 2>&1; tail -5 gpurun_out/r02/gputests_full.log
timeout 600 python bench.py > gpurun_out/r02/bench_n1.json 2> gpurun_out/r02/bench_n1.err; tail -c 600 gpurun_out/r02/bench_n1.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | tail -c 600; echo
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
