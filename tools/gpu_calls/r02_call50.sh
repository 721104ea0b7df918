timeout 100 python tools/check_outer.py 16384 512
echo "== CQR_CHAIN_FUSED=0"; CQR_CHAIN_FUSED=0 timeout 100 python tools/check_outer.py 16384 512
echo "== CQR_PWS_ROWS=0"; CQR_PWS_ROWS=0 timeout 100 python tools/check_outer.py 16384 512
echo "== CQR_PARTITION=0"; CQR_PARTITION=0 timeout 100 python tools/check_outer.py 16384 512 2
echo "== 12288"; timeout 100 python tools/check_outer.py 12288 512 2
