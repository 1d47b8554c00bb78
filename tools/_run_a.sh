set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest.log 2>&1; echo RC=$? >> gpurun_out/r2_gputest.log
tail -3 gpurun_out/r2_gputest.log
timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-side --lanes 2 > gpurun_out/pd_l2.json 2> gpurun_out/pd_l2.err
timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-side --lanes 3 > gpurun_out/pd_l3.json 2> gpurun_out/pd_l3.err
timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-side --lanes 4 > gpurun_out/pd_l4.json 2> gpurun_out/pd_l4.err
timeout 200 python bench.py --no-cpu-baseline --no-side --lanes 2 > gpurun_out/pd_l2e.json 2> gpurun_out/pd_l2e.err
timeout 200 python bench.py --no-cpu-baseline --no-side --lanes 4 > gpurun_out/pd_l4e.json 2> gpurun_out/pd_l4e.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/pd_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), (d.get('e2e') or {}).get('value'), d['parity_frames_ok'], d['parity_frames_checked'])
    except Exception as e: print(f, 'ERR', e)
P
for c in c3 c2 c1 cL; do
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_$c.csv python tools/one_frame.py $c 2 > gpurun_out/r2_ncu_$c.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_vert|k_horiz|k_cost_tile2" -s 3 -c 3 -o gpurun_out/r2_full_c3 -f python tools/one_frame.py c3 2 > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out/r2_*
