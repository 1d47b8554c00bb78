mkdir -p gpurun_out
for l in 4 6; do
timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-side --lanes $l > gpurun_out/rg_l$l.json 2> gpurun_out/rg_l$l.err
B200SGM_VERT_RING=4 timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-side --lanes $l > gpurun_out/rg_r4_l$l.json 2> gpurun_out/rg_r4_l$l.err
done
B200SGM_VERT_RING=4 timeout 120 python tools/stage_time.py c3 12 2>&1 | tail -1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/rg_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), d['parity_frames_ok'], d['parity_frames_checked'])
    except Exception as e: print(f, 'ERR', e)
P
