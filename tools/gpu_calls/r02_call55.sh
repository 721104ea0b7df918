timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/caqr_steps.py 2>&1 | grep -E "factor|Error|error" | head
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q -k caqr 2>&1 | tail -3
