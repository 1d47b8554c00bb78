mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/rf_gputest.log 2>&1; echo RC=$? >> gpurun_out/rf_gputest.log
tail -4 gpurun_out/rf_gputest.log
(timeout 500 python tools/fuzz_parity.py 300 201 2>&1 | tail -1; timeout 300 python tools/fuzz_parity.py 60 202 wide 2>&1 | tail -1; timeout 300 python tools/fuzz_parity.py 30 203 tall 2>&1 | tail -1; timeout 300 python tools/fuzz_sequence.py 100 2>&1 | tail -1) | tee gpurun_out/r2_fuzz4.txt
for c in c1 c2; do for hl in 1 0; do
B200SGM_VERT_HALO=$hl timeout 200 python bench.py --config $c --no-cpu-baseline --no-e2e --no-side --lanes $( [ $c = c1 ] && echo 8 || echo 4 ) > gpurun_out/rj_${c}_h$hl.json 2> gpurun_out/rj_${c}_h$hl.err
done; done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/rj_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), d['parity_frames_ok'], d['parity_frames_checked'])
    except Exception as e: print(f, 'ERR', e)
P
