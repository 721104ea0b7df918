echo "== outer 256"; timeout 200 python tools/la_bench.py 16384 8192
echo "== outer 512"; LA_OUTER=512 timeout 200 python tools/la_bench.py 16384 8192
echo "== outer 384"; LA_OUTER=384 timeout 200 python tools/la_bench.py 16384 8192
timeout 100 python tools/gemm_bench.py gemm 2>&1 | grep "K=512\|K=256" | grep -v "beta=0.0" | head -6
