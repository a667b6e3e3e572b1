cd /root/repo
timeout 300 python -m pytest tests/test_gpu_embedding.py -m gpu -x -q 2>&1 | tail -2
python tools/dbg/rotary_pair_bench.py 2>&1 | tail -4
python - <<'PY'
import torch
from sparse_vae_b200.core.embedding import Embedding
e = Embedding(32768, 512).cuda(); ids = torch.randint(0, 32768, (16, 4096), device='cuda'); dy = torch.randn(16, 4096, 512, device='cuda')
ref = torch.nn.Embedding(32768, 512).cuda()
for name, m in (('own', e), ('aten', ref)):
    y = m(ids)
    for _ in range(3): m.weight.grad = None; y.backward(dy, retain_graph=True)
    torch.cuda.synchronize(); t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(20): m.weight.grad = None; y.backward(dy, retain_graph=True)
    t1.record(); torch.cuda.synchronize(); print(name, 'embedding backward (incl. sort)', round(t0.elapsed_time(t1) / 20 * 1e3, 1), 'us')
PY
