cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_rotary.py -m gpu -x -q 2>&1 | tail -2
python - <<'PY'
import torch
from sparse_vae_b200.core.linear import colsum
x = torch.randn(65536, 512, device='cuda').to(torch.bfloat16); y = torch.randn(65536, 2048, device='cuda').to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for name, t in (('[65536,512]', x), ('[65536,2048]', y)):
    for _ in range(3): colsum(t)
    torch.cuda.synchronize(); tot = 0
    for _ in range(10):
        flush.zero_(); t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(); colsum(t); t1.record(); torch.cuda.synchronize(); tot += t0.elapsed_time(t1)
    print(name, 'colsum (L2 flushed)', round(tot / 10 * 1e3, 1), 'us', round(t.numel() * 2 / (tot / 10 * 1e-3) / 1e9), 'GB/s')
PY
