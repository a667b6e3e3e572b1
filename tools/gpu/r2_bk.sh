cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py tests/test_gpu_graph_step.py tests/test_gpu_linear.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-bottleneck-leg > gpurun_out/r2bk_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2bk_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('clocks'))
print({k: (v['launches'], round(v['us_per_launch'],1)) for k,v in d['kernels'].items() if k in ('grad_scale','grad_sumsq','radam_step','weight_cast')})
PY
