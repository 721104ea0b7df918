mkdir -p gpurun_out/r02
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02/bench_n8.json 2> gpurun_out/r02/bench_n8.err
tail -c 400 gpurun_out/r02/bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
print({k:v for k,v in d['tsqr'].items() if k not in ('roofline','workload')})
print({k:v for k,v in d['caqr'].items() if k not in ('workload',)})
print(d['batched']['ms_per_step'])
PY
