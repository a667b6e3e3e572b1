"""Per-phase clock64 timeline + event timing of the PERSISTENT forward attention kernel at the C2 shape (debug)."""
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
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)
ncta = B * H * (L // 128)
st = torch.cuda.current_stream().cuda_stream


def run(flags, tl=None):
    desc = _make_desc(cfg, q, k, v, out, flags=flags)
    N.check(N.lib.svae_attn_fwd_debug(ctypes.byref(desc), q.data_ptr(), k.data_ptr(), v.data_ptr(), None, out.data_ptr(),
                                      lse.data_ptr(), None, None if tl is None else tl.data_ptr(), st), 'dbg')


import os
for name, flags, stag in (('one CTA per tile', 0, 0), ('persistent', 2, -1), ('persistent', 2, 2000), ('persistent', 2, 0)):
    os.environ['SVAE_FWD_STAGGER'] = str(stag)
    ts = []
    for it in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(flags); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print(f'{name:>18s} stagger {stag}: us per launch (L2 flushed) {[round(x, 1) for x in ts]}  median {np.median(ts[2:]):.1f}')

tl = torch.zeros(ncta, 5, 8, dtype=torch.int64, device=dev)
for it in range(2):
    tl.zero_()
    flush.zero_()
    run(2, tl)
    torch.cuda.synchronize()
t = tl.cpu().numpy()
G = 148
start_gt, end_gt = t[:G, 1, 1], t[:G, 2, 1]
print('globaltimer: kernel span us', (end_gt.max() - start_gt.min()) / 1e3, ' CTA start spread us', (start_gt.max() - start_gt.min()) / 1e3,
      ' CTA durations us: min', ((end_gt - start_gt) / 1e3).min(), 'max', ((end_gt - start_gt) / 1e3).max(),
      ' cycles/us', np.median((t[:G, 2, 0] - t[:G, 1, 0]) / ((end_gt - start_gt) / 1e3)))
dur = (end_gt - start_gt) / 1e3
print('slowest CTAs:', np.argsort(dur)[-6:].tolist(), 'on SMs', t[np.argsort(dur)[-6:], 1, 2].tolist(), 'fastest:', np.argsort(dur)[:6].tolist(), 'on SMs', t[np.argsort(dur)[:6], 1, 2].tolist())


def phases(role, names):
    d = np.diff(t[:, role, :len(names)], axis=1)
    for i in range(len(names) - 1):
        print(f'   {names[i]:>16s} -> {names[i + 1]:<16s} mean {d[:, i].mean():8.0f}  p10 {np.percentile(d[:, i], 10):8.0f}  p90 {np.percentile(d[:, i], 90):8.0f}')


print('softmax warp 0 of the tile\'s group:'); phases(0, ['wait S start', 'S ready', 'max done', 'P arrived'])
print('epilogue warp 8:'); phases(3, ['wait O start', 'O ready', 'O read', 'tile done'])
print('MMA warp 14:'); phases(4, ['wait QK start', 'QK landed', 'tmem free', 'S issued', 'wait P start', 'P ready', 'PV issued'])
for cta in (0, 77):
    ids = np.arange(cta, ncta, G)
    t0 = t[ids[0], 4, 0]
    print(f'CTA {cta}: {len(ids)} tiles, total {t[ids[-1], 3, 3] - t0} cycles; per-tile completion period:',
          np.diff(t[ids, 3, 3])[:14].tolist())
    for nm, role, col in (('QK landed', 4, 1), ('S issued', 4, 3), ('S ready', 0, 1), ('max done', 0, 2), ('P arrived', 0, 3),
                          ('PV issued', 4, 6), ('O ready', 3, 1), ('O read', 3, 2), ('tile done', 3, 3)):
        print(f'    {nm:>10s}:', (t[ids[:8], role, col] - t0).tolist())
