"""Attention forward + backward time per window size at the C2 shape (debug; not a bench)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import sparse_vae_b200 as sv  # noqa: E402
from sparse_vae_b200 import _native as N  # noqa: E402
from util import make_qkv  # noqa: E402

dev = torch.device('cuda')
B, H, L, Dh = 16, 8, 4096, 64
q, k, v = make_qkv(B, H, L, Dh, torch.bfloat16, dev, seed=3, requires_grad=True)
dout = torch.randn(B, L, H * Dh, device=dev, dtype=torch.bfloat16).unflatten(-1, (H, Dh)).transpose(1, 2)
for w in (1, 2, 4, 6, 8, 10):
    cfg = sv.SparseAttention(window_size=w)
    for _ in range(2):
        cfg(q, k, v).backward(dout)
    torch.cuda.synchronize()
    N.profile_begin()
    for _ in range(3):
        cfg(q, k, v).backward(dout)
    torch.cuda.synchronize()
    prof = N.profile_end()
    print(f'window {w:2d}: ' + '  '.join(f"{kname} {val['ms'] / val['launches'] * 1e3:8.1f} us" for kname, val in sorted(prof.items())))
