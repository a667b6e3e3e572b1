cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --config c4 --steps 20 --warmup 3 > gpurun_out/r2bu_bench_c4_n8.json 2> gpurun_out/r2bu_bench_c4_n8.err; echo "c4 n8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2bu_bench_c4_n8.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('notes'), d.get('clocks'))
PY
