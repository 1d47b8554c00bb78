timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "halo or c2 or c1" 2>&1 | tail -4
