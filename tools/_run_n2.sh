mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_c3.json 2> gpurun_out/r2_bench_n2_c3.err; echo "n2 rc=$?"
tail -c 300 gpurun_out/r2_bench_n2_c3.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err; echo "ref n2 rc=$?"
timeout 600 python -m pytest tests/test_host_adapter.py -m gpu -x -q 2>&1 | tail -3
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_n2_c3.json')); print('N2', d['value'], d['e2e']['value'], d['n_gpus'], d.get('parity_frames_ok'), d.get('e2e_plugin',{}).get('ms_per_frame'), d.get('e2e_plugin',{}).get('match_ms_per_frame'))
print(open('gpurun_out/r2_bench_ref_n2.json').read()[:300])
P
