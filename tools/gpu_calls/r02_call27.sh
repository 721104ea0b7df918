export CQR_PANEL_BENCH_MODES=1
R="256 512 768 1024 1536 2048 3072 4096"
echo "== default"; timeout 120 python tools/panel_bench.py $R
for w in 1 2 4 8; do echo "== pair from 1 row, min wpc $w"; CQR_PANEL_PAIR_MIN_ROWS=1 CQR_PANEL_WB_MIN_WPC=$w timeout 120 python tools/panel_bench.py $R; done
mkdir -p gpurun_out/r02
echo "== accuracy probe"; (PROBE_OUTERS=256 timeout 900 python tools/accuracy_probe.py 8192 16384; echo "--- CQR_PANEL_PAIR=0 (one pivot column per exchange)"; CQR_PANEL_PAIR=0 PROBE_OUTERS=256 PROBE_SIMT_MAX=0 timeout 600 python tools/accuracy_probe.py 8192 16384) > gpurun_out/r02/accuracy_probe.txt 2>&1; cat gpurun_out/r02/accuracy_probe.txt
