mkdir -p gpurun_out
(time python bench.py > gpurun_out/r2_bench_n1_c3.json 2> gpurun_out/r2_bench_n1_c3.err) 2> gpurun_out/r2_bench_time.txt
tail -c 400 gpurun_out/r2_bench_n1_c3.err
(time python bench.py --impl reference > gpurun_out/r2_bench_reference_arm_c3.json 2> gpurun_out/r2_bench_ref.err) 2>> gpurun_out/r2_bench_time.txt
cat gpurun_out/r2_bench_time.txt
python __graft_entry__.py smoke 2>&1 | tail -2
