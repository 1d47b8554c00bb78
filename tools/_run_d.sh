mkdir -p gpurun_out
rm -f gpurun_out/rf_st.txt
for c in c3 cL c2 c1; do
timeout 120 python tools/stage_time.py $c 12 2>&1 | tail -1 >> gpurun_out/rf_st.txt
done
cat gpurun_out/rf_st.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/rf_gputest.log 2>&1; echo RC=$? >> gpurun_out/rf_gputest.log
tail -5 gpurun_out/rf_gputest.log
(timeout 400 python tools/fuzz_parity.py 200 91 2>&1 | tail -1; timeout 200 python tools/fuzz_parity.py 40 92 wide 2>&1 | tail -1; timeout 200 python tools/fuzz_parity.py 20 93 tall 2>&1 | tail -1; timeout 300 python tools/fuzz_sequence.py 80 2>&1 | tail -1) | tee gpurun_out/r2_fuzz2.txt
