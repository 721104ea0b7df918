mkdir -p gpurun_out/r02
CQR_CHAIN_FUSED=2 timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02/racecheck_small_final.log 2>&1; tail -12 gpurun_out/r02/racecheck_small_final.log
CQR_CHAIN_FUSED=2 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02/memcheck_small_final.log 2>&1; tail -6 gpurun_out/r02/memcheck_small_final.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:panel_wb2 --launch-skip 3 -c 1 -o gpurun_out/r02/panel_wb2_rot_8192 python tools/panel_bench.py 8192 > gpurun_out/r02/ncu_wb2rot_a.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:panel_wb2 --launch-skip 3 -c 1 -o gpurun_out/r02/panel_wb2_rot_16384 python tools/panel_bench.py 16384 > gpurun_out/r02/ncu_wb2rot_b.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2300 --csv --log-file gpurun_out/r02/launches_geqrf16384_final.csv python tools/one_geqrf.py 16384 > gpurun_out/r02/ncu_launches_final.log 2>&1
python tools/launch_summary.py gpurun_out/r02/launches_geqrf16384_final.csv > gpurun_out/r02/launches_geqrf16384_final.txt; head -24 gpurun_out/r02/launches_geqrf16384_final.txt
ls -la gpurun_out/r02/*.ncu-rep
