mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/rf_gputest.log 2>&1; echo RC=$? >> gpurun_out/rf_gputest.log
tail -3 gpurun_out/rf_gputest.log
(echo "halo-fit strips:"; timeout 500 python tools/fuzz_parity.py 400 701 2>&1 | tail -1; timeout 300 python tools/fuzz_sequence.py 120 2>&1 | tail -1) | tee gpurun_out/r2_fuzz7.txt
