"""clock64 phase timeline of the two backward attention kernels at the C2 shape (debug; not a bench)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import sparse_vae_b200 as sv  # noqa: E402
from sparse_vae_b200 import _native as N  # noqa: E402
from util import make_qkv  # noqa: E402

dev = torch.device('cuda')
B, H, L, Dh = 16, 8, 4096, 64
cfg = sv.SparseAttention()
q, k, v = make_qkv(B, H, L, Dh, torch.bfloat16, dev, seed=3, requires_grad=True)
dout = torch.randn(B, L, H * Dh, device=dev, dtype=torch.bfloat16).unflatten(-1, (H, Dh)).transpose(1, 2)
ncta = B * H * (L // 128)
tl = torch.zeros(2, ncta, 3, 16, dtype=torch.int64, device=dev)
for it in range(3):
    out = cfg(q, k, v)
    if it == 2:
        N.lib.svae_debug_set_bwd_timeline(tl.data_ptr())
    out.backward(dout)
    torch.cuda.synchronize()
    N.lib.svae_debug_set_bwd_timeline(None)
    q.grad = k.grad = v.grad = None
t = tl.cpu().numpy()


def show(kern, role, names):
    x = t[kern, :, role, :len(names)]
    d = np.diff(x, axis=1)
    for i in range(len(names) - 1):
        print(f'   {names[i]:>22s} -> {names[i + 1]:<22s} mean {d[:, i].mean():8.0f}  p10 {np.percentile(d[:, i], 10):8.0f}  p90 {np.percentile(d[:, i], 90):8.0f}')
    print(f'   lifetime mean {np.mean(x[:, -1] - x[:, 0]):8.0f}')


m_dq = ['start', 'alloc+sync', 'loaded+delta', 'sdp0 ready', 'ds0 arrived', 'sdp1 ready', 'ds1 arrived', 'sdp2 ready', 'ds2 arrived',
        'dq ready', 'dq staged', 'G atomics', 'end']
i_dq = ['start', 'alloc+sync', 'TMA issued', 'loaded', 'ds0 ready', 'c0 issued', 'ds1 ready', 'c1 issued', 'ds2 ready', 'c2 issued']
m_kv = ['start', 'alloc+sync', 'stats staged', 'sdp0 ready', 'p0 arrived', 'sdp1 ready', 'p1 arrived', 'sdp2 ready', 'p2 arrived',
        'sdp3 ready', 'p3 arrived', 'out ready', 'end']
i_kv = ['start', 'alloc+sync', 'TMA issued', 'loaded', 'p0 ready', 'c0 issued', 'p1 ready', 'c1 issued', 'p2 ready', 'c2 issued',
        'p3 ready', 'c3 issued']
print('=== dQ pass, math warp 0'); show(0, 0, m_dq)
print('=== dQ pass, math warp 3'); show(0, 1, m_dq)
print('=== dQ pass, MMA warp'); show(0, 2, i_dq)
print('=== dK/dV pass, math warp 0'); show(1, 0, m_kv)
print('=== dK/dV pass, math warp 3'); show(1, 1, m_kv)
print('=== dK/dV pass, MMA warp'); show(1, 2, i_kv)

x = t[0, :, 0, :]
ok = x[:, 13] > 0
print('=== dQ pass, math warp 0, FIRST slot of chunk 0 (live for every tile but the first of a sequence):')
for a, b2, name in ((3, 13, 'sdp0 ready -> scores in registers (2 x tcgen05.ld.x16 + wait)'), (13, 14, 'pair barrier'),
                   (14, 15, '16 elements of exp / dS math + tcgen05.st issue')):
    d = (x[ok, b2] - x[ok, a])
    print(f'   {name:<70s} mean {d.mean():8.0f}  p10 {np.percentile(d, 10):8.0f}  p90 {np.percentile(d, 90):8.0f}')
