"""Debug driver: cross-attention forward, then backward, with a synchronisation after each."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from sparse_vae_b200.core import cross_attention as xa  # noqa: E402

B, H, nq, Lk = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (2, 8, 64, 512)
dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
kk, vv = (torch.randn(B, Lk, H * 64, generator=g).to(dev, torch.bfloat16) for _ in range(2))
k, v = (t.unflatten(-1, (H, 64)).transpose(1, 2).requires_grad_(True) for t in (kk, vv))
q = torch.randn(B, nq, H * 64, generator=g).to(dev, torch.bfloat16).unflatten(-1, (H, 64)).transpose(1, 2).requires_grad_(True)
dout = torch.randn(B, nq, H * 64, generator=g).to(dev, torch.bfloat16).unflatten(-1, (H, 64)).transpose(1, 2)
out = xa.cross_attention(q, k, v, None)
torch.cuda.synchronize()
ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
print('forward ok, rel err', ((out.float() - ref).abs().max() / ref.abs().max()).item(), flush=True)
out.backward(dout)
torch.cuda.synchronize()
qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
torch.nn.functional.scaled_dot_product_attention(qf, kf, vf).backward(dout.float())
for nm, a, b in (('dq', q.grad, qf.grad), ('dk', k.grad, kf.grad), ('dv', v.grad, vf.grad)):
    print(nm, 'rel err', ((a.float() - b).abs().max() / b.abs().max()).item(), flush=True)
