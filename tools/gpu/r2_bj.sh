cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_data_parallel_nccl.py -m gpu -x -q 2>&1 | tail -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2bj_bench_n2.json 2> gpurun_out/r2bj_bench_n2.err; echo "n2 rc=$?"
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-bottleneck-leg > gpurun_out/r2bj_bench_n1.json 2>/dev/null
python - <<'PY'
import json
for f in ('r2bj_bench_n2','r2bj_bench_n1'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('notes'), d.get('clocks'))
PY
