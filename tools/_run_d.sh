mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/rf_gputest.log 2>&1; echo RC=$? >> gpurun_out/rf_gputest.log
tail -4 gpurun_out/rf_gputest.log
(echo "halo depth 3:"; timeout 500 python tools/fuzz_parity.py 300 301 2>&1 | tail -1; timeout 300 python tools/fuzz_parity.py 80 302 wide 2>&1 | tail -1; timeout 300 python tools/fuzz_sequence.py 100 2>&1 | tail -1) | tee gpurun_out/r2_fuzz5.txt
for c in c2 c1 c3; do timeout 120 python tools/stage_time.py $c 12 2>&1 | tail -1; done
