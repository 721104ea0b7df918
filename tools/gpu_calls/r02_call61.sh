mkdir -p gpurun_out/r02
echo "== chunked upload (default)"; timeout 200 python tools/e2e_timeline.py gpurun_out/r02/tl_e2e_chunked.txt 2>&1 | tail -18
echo "== blocking upload"; CQR_H2D_OVERLAP=0 timeout 200 python tools/e2e_timeline.py gpurun_out/r02/tl_e2e_blocking.txt 2>&1 | tail -9
