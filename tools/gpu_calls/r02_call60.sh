timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/caqr_kb_sweep.py 256 512 1024 2048 2>&1 | grep "world"
