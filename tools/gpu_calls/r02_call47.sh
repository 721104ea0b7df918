CQR_CHAIN_COOP_TRACE=1 timeout 100 python tools/one_geqrf.py 8192 2>&1 | tail -8
