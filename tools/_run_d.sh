timeout 120 python tools/stage_time.py cL 12 2>&1 | tail -1
timeout 120 python tools/stage_time.py c3 12 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/fuzz_parity.py 40 401 wide 2>&1 | tail -1
