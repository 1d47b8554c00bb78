set -x
mkdir -p gpurun_out
P=$PWD/i3dr_stereo_camera-ros_b200
for v in "" _nombar; do
export B200SGM_LIB=$P/libb200sgm$v.so
for c in c3 c2 c1 cL; do
timeout 120 python tools/stage_time.py $c 8 2>&1 | tail -1 | tee -a gpurun_out/rd_st$v.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/rd_test$v.log 2>&1; echo RC=$? >> gpurun_out/rd_test$v.log
tail -3 gpurun_out/rd_test$v.log
done
unset B200SGM_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_vert" -s 1 -c 1 -o gpurun_out/rd_full_vert -f python tools/one_frame.py c3 2 > gpurun_out/rd_ncu_full.log 2>&1
ls -la gpurun_out/
