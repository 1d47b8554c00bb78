mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/rf_gputest.log 2>&1; echo RC=$? >> gpurun_out/rf_gputest.log
tail -3 gpurun_out/rf_gputest.log
(echo "final code:"; timeout 500 python tools/fuzz_parity.py 300 901 2>&1 | tail -1; timeout 300 python tools/fuzz_parity.py 60 902 wide 2>&1 | tail -1; timeout 900 python tools/fuzz_parity.py 4 903 full 2>&1 | tail -1; timeout 300 python tools/fuzz_sequence.py 100 2>&1 | tail -1; timeout 300 python tools/fuzz_bm.py 100 9 2>&1 | tail -1) | tee gpurun_out/r2_fuzz9.txt
bash tools/_run_final.sh > gpurun_out/final.log 2>&1
grep "rc=" gpurun_out/final.log | head -2
