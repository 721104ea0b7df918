mkdir -p gpurun_out/r02
timeout 300 python tools/e2e_timeline.py /tmp/tl_overlap.txt > gpurun_out/r02/e2e_timeline_overlap.txt 2>&1
CQR_H2D_OVERLAP=0 timeout 300 python tools/e2e_timeline.py /tmp/tl_block.txt > gpurun_out/r02/e2e_timeline_blocking.txt 2>&1
cp /tmp/tl_overlap.txt gpurun_out/r02/tl_overlap_raw.txt
cat gpurun_out/r02/e2e_timeline_overlap.txt gpurun_out/r02/e2e_timeline_blocking.txt
