cd /root/repo; mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 30 --warmup 3 > gpurun_out/r2bn_bench_n8.json 2> gpurun_out/r2bn_bench_n8.err; echo "n8 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-bottleneck-leg > gpurun_out/r2bn_bench_n4.json 2> gpurun_out/r2bn_bench_n4.err; echo "n4 rc=$?"
python - <<'PY'
import json
for f in ('r2bn_bench_n8','r2bn_bench_n4'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('notes'), d.get('clocks'))
    except Exception as e: print(f, 'ERR', e); print(open(f'gpurun_out/{f}.err').read()[-1200:])
PY
