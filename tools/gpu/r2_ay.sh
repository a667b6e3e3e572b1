cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gelu.py -m gpu -x -q 2>&1 | tail -2
python - <<'PY'
import torch
from sparse_vae_b200.core.gelu import gelu_forward, gelu_backward
x = torch.randn(65536, 2048, device='cuda', dtype=torch.bfloat16); dy = torch.randn_like(x)
def bench(fn, name, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps): fn()
    t1.record(); torch.cuda.synchronize(); print(f'{name:40s} {t0.elapsed_time(t1) / reps * 1e3:8.1f} us')
bench(lambda: gelu_forward(x), 'own gelu fwd [65536,2048]')
bench(lambda: gelu_backward(dy, x, want_colsum=True), 'own gelu bwd + colsum')
bench(lambda: gelu_backward(dy, x), 'own gelu bwd')
PY
python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2ay_kernel_only.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"attn_bwd1|attn_fwd_persist" -s 4 -c 2 -o gpurun_out/r2ay_attn python bench.py --kernel-only --steps 3 --warmup 2 > gpurun_out/r2ay_ncu.log 2>&1; echo "ncu rc=$?"
