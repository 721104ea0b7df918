mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -x -q -k "reference_format or comparator or cli_qr" 2>&1 | tail -30
