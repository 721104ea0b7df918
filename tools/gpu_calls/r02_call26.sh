# smem-traffic experiment on the tcgen05 GEMM (timing only; exp builds compute wrong values)
echo "== baseline"; timeout 200 python tools/gemm_bench.py gemm 2>&1 | grep -E "K=256|K=64" | grep -v "beta=0.0"
for e in 1 2 3; do echo "== exp$e"; CQR_LIB=cuda-qr_b200/csrc/build/exp$e/libcudaqr_b200.so timeout 200 python tools/gemm_bench.py gemm 2>&1 | grep -E "K=256|K=64" | grep -v "beta=0.0"; done
