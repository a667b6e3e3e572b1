"""Per-phase clock64 timeline of the PERSISTENT forward attention kernel at the C2 shape (debug; not a bench)."""
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
desc = _make_desc(cfg, q, k, v, out, flags=2)   # SVAE_ATTN_PERSISTENT
ncta = B * H * (L // 128)
tl = torch.zeros(ncta, 5, 8, dtype=torch.int64, device=dev)
for it in range(3):
    tl.zero_()
    N.check(N.lib.svae_attn_fwd_debug(ctypes.byref(desc), q.data_ptr(), k.data_ptr(), v.data_ptr(), None, out.data_ptr(),
                                      lse.data_ptr(), None, tl.data_ptr(), torch.cuda.current_stream().cuda_stream), 'dbg')
    torch.cuda.synchronize()
t = tl.cpu().numpy()
G = 148
names = ['wait S start', 'S ready', 'max done', 'P arrived', 'pre O wait', 'O ready', 'tile done']
for w in (0, 3):
    d = np.diff(t[:, w, :7], axis=1)
    print(f'softmax warp {w}: mean cycles per phase')
    for i in range(6):
        print(f'   {names[i]:>14s} -> {names[i + 1]:<14s} mean {d[:, i].mean():8.0f}  p10 {np.percentile(d[:, i], 10):8.0f}  p90 {np.percentile(d[:, i], 90):8.0f}')
    print(f'   tile total {np.mean(t[:, w, 6] - t[:, w, 0]):8.0f}')
mn = ['wait QK start', 'QK landed', 'QK issued', 'wait P start', 'P ready', 'PV issued']
d = np.diff(t[:, 4, :6], axis=1)
print('MMA warp:')
for i in range(5):
    print(f'   {mn[i]:>14s} -> {mn[i + 1]:<14s} mean {d[:, i].mean():8.0f}  p10 {np.percentile(d[:, i], 10):8.0f}  p90 {np.percentile(d[:, i], 90):8.0f}')
# one CTA's sequence of tiles: steady-state period
for cta in (0, 77):
    ids = np.arange(cta, ncta, G)
    done = t[ids, 0, 6]
    print(f'CTA {cta}: tiles {len(ids)}, period between consecutive tile completions (cycles):', np.diff(done)[:12].tolist())
    print('    S-ready times rel:', (t[ids[:8], 0, 1] - t[ids[0], 0, 0]).tolist())
    print('    QK landed rel   :', (t[ids[:8], 4, 1] - t[ids[0], 0, 0]).tolist())
    print('    QK wait start   :', (t[ids[:8], 4, 0] - t[ids[0], 0, 0]).tolist())

for cta in (0, 77):
    ids = np.arange(cta, ncta, G)
    t0 = t[ids[0], 0, 0]
    print(f'CTA {cta} producers (rel cycles), tiles 0..7:')
    print('    QK  wait-start:', (t[ids[:8], 1, 0] - t0).tolist())
    print('    QK  freed     :', (t[ids[:8], 1, 1] - t0).tolist())
    print('    QK  issued    :', (t[ids[:8], 1, 2] - t0).tolist())
    print('    QK  landed    :', (t[ids[:8], 4, 1] - t0).tolist())
    print('    QK  mma issued:', (t[ids[:8], 4, 2] - t0).tolist())
    print('    S ready (w0)  :', (t[ids[:8], 0, 1] - t0).tolist())
    print('    V   freed     :', (t[ids[:8], 2, 1] - t0).tolist())
    print('    V   issued    :', (t[ids[:8], 2, 2] - t0).tolist())
    print('    P arrived w0  :', (t[ids[:8], 0, 3] - t0).tolist())
    print('    PV wait start :', (t[ids[:8], 4, 3] - t0).tolist())
    print('    PV P ready    :', (t[ids[:8], 4, 4] - t0).tolist())
    print('    PV issued     :', (t[ids[:8], 4, 5] - t0).tolist())
    print('    O ready (w0)  :', (t[ids[:8], 0, 5] - t0).tolist())
    print('    tile done (w0):', (t[ids[:8], 0, 6] - t0).tolist())
