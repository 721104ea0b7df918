mkdir -p gpurun_out/r02
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r02/bench_n4.json 2> gpurun_out/r02/bench_n4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_n4.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
t=d['tsqr']; print('tsqr', t['ms_per_step'], t.get('local_ms'), t.get('cross_gpu_ms'), t.get('efficiency_vs_ideal'))
c=d['caqr']; print('caqr', c['ms_per_step'], c['value'], c.get('backward_error_over_n_eps'), c.get('orthogonality_over_n_eps'))
PY
