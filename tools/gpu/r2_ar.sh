cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_ce.py -m gpu -x -q 2>&1 | tail -6
python - <<'PY'
import torch
from sparse_vae_b200 import _native as N
rows, V = 16384, 32768
x = (torch.randn(rows, V, device='cuda') * 2).to(torch.bfloat16)
labels = torch.randint(1, V, (rows,), device='cuda'); w = torch.full((rows,), 1.0 / rows, device='cuda'); nll = torch.empty(rows, device='cuda')
for variant, name in ((1, 'streamed'), (3, 'register-resident')):
    buf = x.clone()
    def run(): N.check(N.lib.svae_vocab_ce(buf.data_ptr(), N.DTYPE_BF16, rows, V, V, labels.data_ptr(), w.data_ptr(), nll.data_ptr(), variant, N.current_stream(buf.device)), 'ce')
    for _ in range(3): run()
    torch.cuda.synchronize(); t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10): run()
    t1.record(); torch.cuda.synchronize()
    us = t0.elapsed_time(t1) / 10 * 1e3
    print(f'{name:20s} {us:8.1f} us  {2 * rows * V * 2 / us / 1e3:7.1f} GB/s')
PY
