"""Runs the forward attention kernel a few times at the C2 shape (for ncu).  Usage: run_fwd_once.py [persistent=0|1]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
os.environ['SVAE_ATTN_PERSISTENT'] = sys.argv[1] if len(sys.argv) > 1 else '1'
import sparse_vae_b200 as sv  # noqa: E402
from util import make_qkv  # noqa: E402

q, k, v = make_qkv(16, 8, 4096, 64, torch.bfloat16, torch.device('cuda'), seed=3)
cfg = sv.SparseAttention()
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device='cuda')
for _ in range(3):
    flush.zero_()
    out = cfg(q, k, v)
torch.cuda.synchronize()
print('ok', float(out.float().abs().mean()))
