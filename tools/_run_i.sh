mkdir -p gpurun_out
for c in c2 c1; do for l in 4 8; do
timeout 200 python bench.py --config $c --no-cpu-baseline --no-e2e --no-side --lanes $l > gpurun_out/ri_${c}_l$l.json 2> gpurun_out/ri_${c}_l$l.err
done; done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/ri_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), d['parity_frames_ok'], d['parity_frames_checked'])
    except Exception as e: print(f, 'ERR', e)
P
