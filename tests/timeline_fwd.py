"""Per-phase clock64 timeline of the forward attention kernel at the C2 shape (debug entry point; not a bench)."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import sparse_vae_b200 as sv  # noqa: E402
from sparse_vae_b200 import _native as N  # noqa: E402
from sparse_vae_b200.core.sparse_attention import _make_desc, _new_blhd  # noqa: E402
from util import make_qkv  # noqa: E402

dev = torch.device('cuda')
B, H, L, Dh = 16, 8, 4096, 64
cfg = sv.SparseAttention()
q, k, v = make_qkv(B, H, L, Dh, torch.bfloat16, dev, seed=3)
out = _new_blhd(B, H, L, Dh, q)
lse = torch.empty(B, H, L, device=dev)
desc = _make_desc(cfg, q, k, v, out)
ncta = B * H * (L // 128)
tl = torch.zeros(ncta, 5, 8, dtype=torch.int64, device=dev)
for it in range(3):
    tl.zero_()
    N.check(N.lib.svae_attn_fwd_debug(ctypes.byref(desc), q.data_ptr(), k.data_ptr(), v.data_ptr(), None, out.data_ptr(),
                                      lse.data_ptr(), None, tl.data_ptr(), torch.cuda.current_stream().cuda_stream), 'dbg')
    torch.cuda.synchronize()
t = tl.cpu().numpy()
sm = t[:, 4, 7] & 255
gt = t[:, 4, 7] >> 8
print('kernel span (globaltimer, us):', (gt.max() - gt.min()) / 1e3)
names_s = ['start', 'alloc+sync', 'S ready', 'pass1 done', 'pass2 done(arrive)', 'O ready', 'epilogue done']
names_m = ['start', 'alloc+sync', 'TMA issued', 'QK landed', 'QK mma issued', 'P ready', 'PV issued']
for w, names in ((0, names_s), (3, names_s), (4, names_m)):
    d = np.diff(t[:, w, :7], axis=1)
    print(f'warp {w}: mean cycles per phase')
    for i in range(6):
        print(f'   {names[i]:>20s} -> {names[i + 1]:<20s} mean {d[:, i].mean():8.0f}  p10 {np.percentile(d[:, i], 10):8.0f}  p90 {np.percentile(d[:, i], 90):8.0f}')
    print(f'   total {np.mean(t[:, w, 6] - t[:, w, 0]):8.0f}')
# per-SM concurrency: CTAs per SM and busy fraction
life = t[:, 0, 6] - t[:, 0, 0]
print('CTA lifetime cycles mean', life.mean(), 'CTAs per SM mean', np.bincount(sm).mean(), 'SMs used', len(np.unique(sm)))
for s in np.unique(sm)[:2]:
    idx = np.where(sm == s)[0]
    order = idx[np.argsort(t[idx, 4, 0])]
    print('SM', s, 'first CTA starts (relative cycles):', (t[order[:8], 4, 0] - t[order[0], 4, 0]).tolist())
    print('          ends:', (t[order[:8], 0, 6] - t[order[0], 4, 0]).tolist())
