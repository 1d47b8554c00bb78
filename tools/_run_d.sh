timeout 120 python tools/stage_time.py cL 12 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cL or wide_disparity" 2>&1 | tail -2
