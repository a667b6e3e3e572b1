cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_gelu.py tests/test_gpu_model.py tests/test_gpu_graph_step.py -m gpu -x -q 2>&1 | tail -6
python bench.py --kernel-only --steps 20 --warmup 3 > gpurun_out/r2ap_kernel_only.json 2> gpurun_out/r2ap_kernel_only.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ap_kernel_only.json'))
print({k: round(x['us_per_launch'],1) for k,x in d['kernels'].items()})
PY
python bench.py --steps 30 --warmup 3 > gpurun_out/r2ap_bench.json 2> gpurun_out/r2ap_bench.err; tail -2 gpurun_out/r2ap_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ap_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('notes'))
print({k: round(v['us_per_launch'],1) for k,v in d['kernels'].items()})
PY
