# ncu evidence for the Gram leaf (gram_umma.cu), 8M x 64
set -x
mkdir -p gpurun_out/r02b
CQR_TSQR_LEAF=gram timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b/launches_tsqr_gram.csv \
  python tools/tsqr_bench.py once 8388608 > gpurun_out/r02b/ncu_gram_l.log 2>&1
tail -12 gpurun_out/r02b/launches_tsqr_gram.csv
CQR_TSQR_LEAF=gram timeout 600 ncu --set full --import-source on --clock-control none -k regex:gram_kernel -c 1 -o gpurun_out/r02b/tsqr_gram \
  python tools/tsqr_bench.py once 8388608 > gpurun_out/r02b/ncu_gram.log 2>&1
ls -la gpurun_out/r02b/
