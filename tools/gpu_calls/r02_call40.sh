mkdir -p gpurun_out/r02
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02/bench_n8.json 2> gpurun_out/r02/bench_n8.err
tail -c 3000 gpurun_out/r02/bench_n8.json; tail -5 gpurun_out/r02/bench_n8.err
