mkdir -p gpurun_out/r02
echo "== defaults"; timeout 300 python tools/size_sweep.py | tee gpurun_out/r02/size_sweep_default.txt
echo "== three-launch updates, slices from 14336 rows (start of the day)"; CQR_CHAIN_FUSED=0 CQR_PWS_ROWS=14336 timeout 300 python tools/size_sweep.py | tee gpurun_out/r02/size_sweep_old.txt
