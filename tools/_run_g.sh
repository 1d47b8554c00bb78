timeout 600 python -m pytest tests/test_gpu_bm.py tests/test_host_adapter.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/fuzz_bm.py 150 2>&1 | tail -3
