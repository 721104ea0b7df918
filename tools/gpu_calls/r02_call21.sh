mkdir -p gpurun_out/r02
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/caqr_kb_sweep.py 256 512 1024 2048 > gpurun_out/r02/caqr_kb_sweep_n2.txt 2>&1
grep "world" gpurun_out/r02/caqr_kb_sweep_n2.txt
