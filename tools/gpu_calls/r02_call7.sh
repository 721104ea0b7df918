mkdir -p gpurun_out/r02
run() { echo "== $*"; env "$@" timeout 200 python bench.py --no-extra --no-cpu --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('device', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
{
run CQR_X=0
run CQR_H2D_GBPS=55
run CQR_H2D_GBPS=55 CQR_H2D_BLOCK_MS=1.3
run CQR_H2D_GBPS=45 CQR_H2D_BLOCK_MS=0.8
run CQR_CATCH_COLS=1024
run CQR_CATCH_COLS=512
run CQR_CATCH_CTAS_PCT=50
run CQR_CATCH_CTAS_PCT=50 CQR_CATCH_COLS=1024
run CQR_H2D_GBPS=55 CQR_H2D_BLOCK_MS=1.3 CQR_CATCH_COLS=1024
run CQR_PARTITION=0
} > gpurun_out/r02/e2e_sweep.txt 2>&1
cat gpurun_out/r02/e2e_sweep.txt
python - <<'PY'
import torch, time
n=1<<28
h=torch.empty(n, dtype=torch.float32, pin_memory=True); d=torch.empty(n, device='cuda')
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); t1=time.perf_counter()-t
    torch.cuda.synchronize(); t=time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); t2=time.perf_counter()-t
print('H2D GB/s', n*4/t1/1e9, 'D2H GB/s', n*4/t2/1e9)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream(); h2=torch.empty(n, dtype=torch.float32, pin_memory=True); d2=torch.empty(n, device='cuda')
torch.cuda.synchronize(); t=time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); t3=time.perf_counter()-t
print('both directions at once: GB/s each', n*4/t3/1e9)
PY
