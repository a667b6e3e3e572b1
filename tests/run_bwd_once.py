"""Debug driver for the one-pass backward: runs a few shapes, compares dq / dk / dv with the exact CUDA-core path and
the two-pass tcgen05 kernels on the device and prints the per-128-row-tile error profile of the worst (batch, head).

    python tests/run_bwd_once.py [B H L window [cls]] ...
"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import sparse_vae_b200 as sv  # noqa: E402
from util import make_padding, make_qkv  # noqa: E402


def run(B, H, L, window, cls=True, lengths=None, seed=0):
    dev = torch.device('cuda')
    cfg = sv.SparseAttention(window_size=window, include_cls=cls, num_heads=H)
    q, k, v = make_qkv(B, H, L, 64, torch.bfloat16, dev, seed=seed, requires_grad=True)
    g = torch.Generator().manual_seed(seed + 1)
    dout = torch.randn(B, L, H * 64, generator=g).to(dev, torch.bfloat16).unflatten(-1, (H, 64)).transpose(1, 2)
    kpm = make_padding(B, L, lengths, dev) * -1e7 if lengths else None
    res = {}
    for name, env, exact in (('one_pass', '0', False), ('two_pass', '1', False), ('exact', '0', True)):
        os.environ['SVAE_ATTN_BWD_TWO_PASS'] = env
        q.grad = k.grad = v.grad = None
        out = cfg(q, k, v, key_padding_mask=kpm, force_exact=exact)
        out.backward(dout)
        torch.cuda.synchronize()
        res[name] = [t.grad.float().clone() for t in (q, k, v)]
    os.environ['SVAE_ATTN_BWD_TWO_PASS'] = '0'
    print(f'B={B} H={H} L={L} window={window} cls={cls} lengths={lengths}')
    for i, nm in enumerate(('dq', 'dk', 'dv')):
        ref = res['exact'][i]
        scale = ref.abs().max().item()
        line = f'  {nm}: |ref|max {scale:.3e}'
        for other in ('one_pass', 'two_pass'):
            a = res[other][i]
            nan = torch.isnan(a) ^ torch.isnan(ref)
            err = (torch.nan_to_num(a) - torch.nan_to_num(ref)).abs()
            line += f' | {other} rel {err.max().item() / scale:.3e} nan-mismatch {int(nan.sum())}'
        print(line)
        a = res['one_pass'][i]
        err = (torch.nan_to_num(a) - torch.nan_to_num(ref)).abs() / scale
        if err.max().item() > 1e-2:
            b, h = divmod(int(err.flatten(2).amax(-1).argmax()), H)
            nt = (L + 127) // 128
            pad = nt * 128 - L
            e = torch.nn.functional.pad(err[b, h], (0, 0, 0, pad)).reshape(nt, 4, 32 * 64).amax(-1)
            print(f'    worst (b={b}, h={h}); per tile x 32-row block max rel err:')
            for t in range(nt):
                print('     tile %3d: %s' % (t, ' '.join(f'{x:.1e}' for x in e[t].tolist())))


if __name__ == '__main__':
    args = sys.argv[1:]
    if args:
        vals = [int(a) for a in args]
        run(vals[0], vals[1], vals[2], vals[3], bool(vals[4]) if len(vals) > 4 else True)
    else:
        run(1, 1, 128, 4)
        run(1, 1, 256, 4)
        run(1, 2, 512, 4)
        run(2, 8, 1024, 4, lengths=[1024, 700])
        run(1, 8, 640, 2)
        run(1, 8, 640, 1)
        run(1, 4, 608, 3, cls=False)
        run(5, 8, 4096, 4)
        run(1, 2, 16384, 4)
