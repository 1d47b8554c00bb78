mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/rl_bench.json 2> gpurun_out/rl_bench.err; echo rc=$?
tail -c 300 gpurun_out/rl_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/rl_bench.json'))
print(d['value'], d['parity_frames_ok'])
for k,v in d['configs'].items(): print(k, v.get('frames_per_s'), v.get('single_lane_ms'), v.get('one_lane_engine_ms'), v.get('one_lane_engine_ok'), v.get('parity_frames_ok'))
P
