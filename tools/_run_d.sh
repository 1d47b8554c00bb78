mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/rf_gputest.log 2>&1; echo RC=$? >> gpurun_out/rf_gputest.log
tail -4 gpurun_out/rf_gputest.log
(echo "halo producers, depth 4:"; timeout 500 python tools/fuzz_parity.py 400 601 2>&1 | tail -1; timeout 300 python tools/fuzz_parity.py 120 602 wide 2>&1 | tail -1; timeout 300 python tools/fuzz_parity.py 20 603 tall 2>&1 | tail -1; timeout 300 python tools/fuzz_sequence.py 120 2>&1 | tail -1) | tee gpurun_out/r2_fuzz6.txt
for c in c2 c1 c3 cL; do timeout 120 python tools/stage_time.py $c 12 2>&1 | tail -1; done
