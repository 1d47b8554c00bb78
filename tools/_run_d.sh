mkdir -p gpurun_out
rm -f gpurun_out/rf_st.txt
for rep in 1 2; do
for c in c3 cL c2 c1; do
timeout 120 python tools/stage_time.py $c 12 2>&1 | tail -1 >> gpurun_out/rf_st.txt
done
done
cat gpurun_out/rf_st.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/rf_gputest.log 2>&1; echo RC=$? >> gpurun_out/rf_gputest.log
tail -3 gpurun_out/rf_gputest.log
