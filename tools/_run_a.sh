for c in c1 c2 c3 cL; do python tools/stage_time.py $c 6 2>&1 | tail -1; done
for c in c1 c2; do B200SGM_VERT_ALL_SMS=1 python tools/stage_time.py $c 6 2>&1 | tail -1; done
for c in c1 c2 c3; do
python bench.py --config $c --no-cpu-baseline --no-e2e --no-side > gpurun_out/pb_$c.json 2> gpurun_out/pb_$c.err
B200SGM_VERT_ALL_SMS=1 python bench.py --config $c --no-cpu-baseline --no-e2e --no-side > gpurun_out/pb_${c}_all.json 2> gpurun_out/pb_${c}_all.err
B200SGM_VERT_PLAIN=1 python bench.py --config $c --no-cpu-baseline --no-e2e --no-side > gpurun_out/pb_${c}_plain.json 2> gpurun_out/pb_${c}_plain.err
done
python bench.py --config c1 --lanes 8 --no-cpu-baseline --no-e2e --no-side > gpurun_out/pb_c1_l8.json 2> gpurun_out/pb_c1_l8.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/pb_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), d['parity_frames_ok'], d['parity_frames_checked'])
    except Exception as e: print(f, 'ERR', e)
P
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
