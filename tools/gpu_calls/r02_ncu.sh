# ncu evidence of round 2 (1 GPU).  Green contexts switch themselves off under the injecting profiler.
set -x
mkdir -p gpurun_out/r02
# (1) launch list of the second 16384^2 factorisation of a bench run (device-resident step only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 10070 -c 10200 --csv --log-file gpurun_out/r02/launches_bench16384.csv \
  python bench.py --no-extra --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/r02/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r02/launches_bench16384.csv > gpurun_out/r02/launches_bench16384.txt; head -20 gpurun_out/r02/launches_bench16384.txt
# (2) --set full: two-column panel kernel, one cluster (8192 rows) and two clusters (16384 rows)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:panel_wb2 --launch-skip 3 -c 1 -o gpurun_out/r02/panel_wb2_8192 python tools/panel_bench.py 8192 > gpurun_out/r02/ncu_wb2a.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:panel_wb2 --launch-skip 3 -c 1 -o gpurun_out/r02/panel_wb2_16384 python tools/panel_bench.py 16384 > gpurun_out/r02/ncu_wb2b.log 2>&1
# (3) --set full: the two trailing-update GEMMs
timeout 600 ncu --set full --import-source on --clock-control none -k regex:umma_gemm -c 2 -o gpurun_out/r02/gemm python tools/gemm_bench.py gemm once > gpurun_out/r02/ncu_gemm.log 2>&1
# (4) --set full: TSQR leaves (SIMT flat tree and tensor-pipe flat tree), batched warp kernel
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tsqr_flat_r -c 1 -o gpurun_out/r02/tsqr_flat python tools/tsqr_bench.py once 8388608 > gpurun_out/r02/ncu_flat.log 2>&1
CQR_TSQR_LEAF=mma timeout 600 ncu --set full --import-source on --clock-control none -k regex:tsqr_mma_r -c 1 -o gpurun_out/r02/tsqr_mma python tools/tsqr_bench.py once 8388608 > gpurun_out/r02/ncu_mma.log 2>&1
ls -la gpurun_out/r02/*.ncu-rep
