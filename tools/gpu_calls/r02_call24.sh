# ncu --set full of the fused chain update (forced on: kernels are serialised under ncu), first launches of a 16384^2 run (m_p ~ 16320)
mkdir -p gpurun_out/r02
CQR_CHAIN_FUSED=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:chain_update --launch-skip 4 -c 1 -o gpurun_out/r02/chain_update python tools/one_geqrf.py 16384 > gpurun_out/r02/ncu_chain.log 2>&1
tail -5 gpurun_out/r02/ncu_chain.log
ls -la gpurun_out/r02/chain_update.ncu-rep
