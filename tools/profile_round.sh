# One round's profile set on a B200 (run through gpurun): both bench arms, ncu launch lists of one frame per config, full captures of
# the volume kernels at c3 and cL.  Post-process here with tools/summarize_launches.py, `ncu -i ... --page details` and tools/ncu_stalls.py.
set -x
mkdir -p gpurun_out
# 1. bench lines (both arms), default flags
python bench.py > gpurun_out/r2_bench_n1_c3.json 2> gpurun_out/r2_bench_n1_c3.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm_c3.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
# 2. ncu launch lists of one frame per config (after the programs above exited 0 without ncu)
for c in c3 c2 c1 cL; do
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_$c.csv python tools/one_frame.py $c 2 > gpurun_out/r2_ncu_$c.log 2>&1
done
# 3. full captures of the three volume kernels at c3 and of the cL ones
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_vert|k_horiz|k_cost_tile2" -s 3 -c 3 -o gpurun_out/r2_full_c3 -f python tools/one_frame.py c3 2 > gpurun_out/r2_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_vert|k_horiz|k_cost_tile2" -s 3 -c 3 -o gpurun_out/r2_full_cL -f python tools/one_frame.py cL 2 > gpurun_out/r2_ncu_full_cL.log 2>&1
ls -la gpurun_out/
