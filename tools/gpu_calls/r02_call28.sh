export CQR_LIB=cuda-qr_b200/csrc/build/trace/libcudaqr_b200.so
CQR_PANEL_PAIR_MIN_ROWS=1 timeout 200 python tools/wb2_trace.py 256 512 2048 8192 16384
