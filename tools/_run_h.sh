mkdir -p gpurun_out
(timeout 500 python tools/fuzz_parity.py 250 77 2>&1 | tail -2
timeout 300 python tools/fuzz_parity.py 60 78 wide 2>&1 | tail -2
timeout 300 python tools/fuzz_parity.py 30 79 tall 2>&1 | tail -2
timeout 300 python tools/fuzz_sequence.py 100 2>&1 | tail -2) | tee gpurun_out/r2_fuzz.txt
