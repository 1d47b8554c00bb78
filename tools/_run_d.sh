mkdir -p gpurun_out
rm -f gpurun_out/rf_st.txt
P=$PWD/i3dr_stereo_camera-ros_b200
for c in c3 cL c2 c1; do
timeout 120 python tools/stage_time.py $c 12 2>&1 | tail -1 >> gpurun_out/rf_st.txt
B200SGM_LIB=$P/libb200sgm_split.so timeout 120 python tools/stage_time.py $c 12 2>&1 | tail -1 >> gpurun_out/rf_st.txt
done
cat gpurun_out/rf_st.txt
B200SGM_LIB=$P/libb200sgm_split.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/rf_gputest.log 2>&1; echo RC=$? >> gpurun_out/rf_gputest.log
tail -5 gpurun_out/rf_gputest.log
