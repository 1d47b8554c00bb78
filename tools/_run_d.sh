mkdir -p gpurun_out
rm -f gpurun_out/rf_st.txt
for c in c3 c2 c1 cL; do
timeout 120 python tools/stage_time.py $c 12 2>&1 | tail -1 >> gpurun_out/rf_st.txt
done
B200SGM_VERT_STAGE8=0 timeout 120 python tools/stage_time.py c3 12 2>&1 | tail -1 >> gpurun_out/rf_st.txt
cat gpurun_out/rf_st.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/rf_gputest.log 2>&1; echo RC=$? >> gpurun_out/rf_gputest.log
tail -5 gpurun_out/rf_gputest.log
