mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests -m gpu -x -q -k "one_launch_chain_update" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02/bench_n2.json 2> gpurun_out/r02/bench_n2.err
tail -c 400 gpurun_out/r02/bench_n2.json
