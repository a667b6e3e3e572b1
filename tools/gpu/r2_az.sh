cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_data_parallel_nccl.py -m gpu -x -q 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2az_bench_n8.json 2> gpurun_out/r2az_bench_n8.err; echo "n8 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --config c4 --steps 10 --warmup 3 > gpurun_out/r2az_bench_c4_n8.json 2> gpurun_out/r2az_bench_c4_n8.err; echo "c4 n8 rc=$?"
python - <<'PY'
import json
for f in ('r2az_bench_n8','r2az_bench_c4_n8'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('notes'), d.get('clocks'))
    except Exception as e: print(f, 'ERR', e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY
