mkdir -p gpurun_out
timeout 120 python tools/stage_time.py cL 12 2>&1 | tail -1
timeout 120 python tools/stage_time.py c3 12 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/rf_gputest.log 2>&1; echo RC=$? >> gpurun_out/rf_gputest.log
tail -3 gpurun_out/rf_gputest.log
(echo "k_horiz: unrolled blocks for padded disparity counts:"; timeout 500 python tools/fuzz_parity.py 400 801 2>&1 | tail -1; timeout 300 python tools/fuzz_parity.py 80 802 wide 2>&1 | tail -1; timeout 300 python tools/fuzz_parity.py 20 803 tall 2>&1 | tail -1) | tee gpurun_out/r2_fuzz8.txt
