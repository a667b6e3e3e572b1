cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rotary.py tests/test_gpu_attention.py tests/test_gpu_model.py tests/test_gpu_graph_step.py tests/test_gpu_training_curve.py tests/test_gpu_linear.py tests/test_gpu_compat_trainer.py -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 30 --warmup 3 > gpurun_out/r2bg_bench.json 2> gpurun_out/r2bg_bench.err; tail -2 gpurun_out/r2bg_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2bg_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('notes'), d.get('clocks'), d.get('final_loss'))
print({k: (v['launches'], round(v['us_per_launch'],1)) for k,v in d['kernels'].items()})
PY
