cd /root/repo; mkdir -p gpurun_out
python tools/dbg/test_norm_probe.py 2>&1 | grep "^\[" 
python -m pytest tools/dbg/test_norm_probe.py -q -s 2>&1 | grep "^\["
timeout 600 python -m pytest tests/test_gpu_gelu.py tests/test_gpu_linear.py tests/test_gpu_model.py tests/test_gpu_graph_step.py -m gpu -x -q 2>&1 | tail -8
python - <<'PY'
import torch, time
from sparse_vae_b200.core.gelu import gelu_forward, gelu_backward
x = torch.randn(65536, 2048, device='cuda', dtype=torch.bfloat16); dy = torch.randn_like(x)
def bench(fn, name, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps): fn()
    t1.record(); torch.cuda.synchronize(); print(f'{name:40s} {t0.elapsed_time(t1) / reps * 1e3:8.1f} us')
bench(lambda: gelu_forward(x), 'own gelu fwd [65536,2048]')
bench(lambda: torch.nn.functional.gelu(x), 'aten gelu fwd')
bench(lambda: gelu_backward(dy, x, want_colsum=True), 'own gelu bwd + colsum')
bench(lambda: gelu_backward(dy, x), 'own gelu bwd')
bench(lambda: torch.ops.aten.gelu_backward(dy, x), 'aten gelu bwd')
PY
python bench.py --steps 30 --warmup 3 > gpurun_out/r2ao_bench.json 2> gpurun_out/r2ao_bench.err; tail -2 gpurun_out/r2ao_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ao_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('notes'))
print({k: round(v['us_per_launch'],1) for k,v in d['kernels'].items()})
PY
