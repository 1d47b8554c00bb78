set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gputest.log 2>&1; echo RC=$? >> gpurun_out/r2b_gputest.log
tail -3 gpurun_out/r2b_gputest.log
for c in c3 c2 c1 cL; do
timeout 120 python tools/stage_time.py $c 8 2>&1 | tail -1 | tee -a gpurun_out/st_base.txt
B200SGM_LIB=$PWD/i3dr_stereo_camera-ros_b200/libb200sgm_cps2.so timeout 120 python tools/stage_time.py $c 8 2>&1 | tail -1 | tee -a gpurun_out/st_cps2.txt
done
timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-side --lanes 4 > gpurun_out/b_base.json 2> gpurun_out/b_base.err
B200SGM_LIB=$PWD/i3dr_stereo_camera-ros_b200/libb200sgm_cps2.so timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-side --lanes 4 > gpurun_out/b_cps2.json 2> gpurun_out/b_cps2.err
tail -c 600 gpurun_out/b_base.json; tail -c 600 gpurun_out/b_cps2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_vert" -s 1 -c 1 -o gpurun_out/r2b_full_vert -f python tools/one_frame.py c3 2 > gpurun_out/r2b_ncu_full.log 2>&1
B200SGM_LIB=$PWD/i3dr_stereo_camera-ros_b200/libb200sgm_cps2.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_vert" -s 1 -c 1 -o gpurun_out/r2b_full_vert_cps2 -f python tools/one_frame.py c3 2 > gpurun_out/r2b_ncu_full_cps2.log 2>&1
ls -la gpurun_out/
