mkdir -p gpurun_out/r02
nvidia-smi topo -m 2>&1 | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/dist_tsqr_bench.py 8388608 1048576 > gpurun_out/r02/dist_tsqr_bench_n2.txt 2>&1
cat gpurun_out/r02/dist_tsqr_bench_n2.txt | tail -8
