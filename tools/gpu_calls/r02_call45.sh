echo "== pws by work"; timeout 300 python tools/size_sweep.py
