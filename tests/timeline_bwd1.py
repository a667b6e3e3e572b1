"""Where the one-pass backward spends its time: runs svae_attn_bwd from libsvae_b200_dbg.so (the product kernels
compiled with -DSVAE_DEBUG_BUILD) at the C2 shape and prints, per warp role, the cycles spent in every kind of wait
(mean over CTAs).  Build the library with `python sparse_vae_b200/csrc/build.py --debug`."""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import sparse_vae_b200 as sv  # noqa: E402
from sparse_vae_b200 import _native as N  # noqa: E402
from sparse_vae_b200.core.sparse_attention import _make_desc, _new_blhd, _strides3  # noqa: E402
from util import make_qkv  # noqa: E402

KINDS = ['full', 'stat', 's_ready', 'p_ready', 'u_free', 'group', 'ds_free', 'acc_ready', 'acc_free', 'free']
ROLES = {0: 'math0.q0', 1: 'math0.q1', 3: 'math0.q3', 4: 'math1.q0', 7: 'math1.q3', 8: 'epilogue.q0', 12: 'tma', 13: 'S/dP issue',
         14: 'dV/dK issue', 15: 'dQ/G issue'}


def main():
    B, H, L = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (16, 8, 4096)
    dbg = N.load_debug()
    dbg.svae_attn_bwd.restype = ctypes.c_int
    dbg.svae_attn_bwd.argtypes = N.lib.svae_attn_bwd.argtypes
    dbg.svae_attn_bwd_workspace_bytes.restype = ctypes.c_size_t
    dbg.svae_attn_bwd_workspace_bytes.argtypes = N.lib.svae_attn_bwd_workspace_bytes.argtypes
    dbg.svae_debug_set_b1_timeline.restype = None
    dbg.svae_debug_set_b1_timeline.argtypes = [ctypes.c_void_p]
    dbg.svae_debug_set_b1_knock.restype = None
    dbg.svae_debug_set_b1_knock.argtypes = [ctypes.c_int]
    dbg.svae_last_error.restype = ctypes.c_char_p
    dev = torch.device('cuda')
    cfg = sv.SparseAttention()
    q, k, v = make_qkv(B, H, L, 64, torch.bfloat16, dev, seed=1)
    dout = torch.randn(B, L, H * 64, device=dev, dtype=torch.bfloat16).unflatten(-1, (H, 64)).transpose(1, 2)
    with torch.no_grad():
        out = cfg(q, k, v)
    # lse is not returned by the public call: recompute it through the forward entry point
    lse = torch.empty(B, H, L, device=dev)
    out2 = _new_blhd(B, H, L, 64, q)
    desc = _make_desc(cfg, q, k, v, out2, N.ATTN_PERSISTENT)
    st = torch.cuda.current_stream().cuda_stream
    N.check(N.lib.svae_attn_fwd(ctypes.byref(desc), q.data_ptr(), k.data_ptr(), v.data_ptr(), None, out2.data_ptr(), lse.data_ptr(), st), 'fwd')
    dq, dk, dv = (_new_blhd(B, H, L, 64, q) for _ in range(3))
    desc.do_stride, desc.dq_stride, desc.dk_stride, desc.dv_stride = _strides3(dout), _strides3(dq), _strides3(dk), _strides3(dv)
    ws_bytes = dbg.svae_attn_bwd_workspace_bytes(ctypes.byref(desc))
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
    ws_ptr = (ws.data_ptr() + 255) & ~255
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    tl = torch.zeros(sms, 16, 16, dtype=torch.int64, device=dev)
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)
    def with_timeline(mask):
        dbg.svae_debug_set_b1_knock(mask)
        tl.zero_()
        for it in range(3):
            flush.zero_()
            dbg.svae_debug_set_b1_timeline(tl.data_ptr() if it == 2 else None)
            rc = dbg.svae_attn_bwd(ctypes.byref(desc), q.data_ptr(), k.data_ptr(), v.data_ptr(), out2.data_ptr(), dout.data_ptr(),
                                   lse.data_ptr(), None, dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ws_ptr, ws_bytes, st)
            assert rc == 0, dbg.svae_last_error()
            torch.cuda.synchronize()
        dbg.svae_debug_set_b1_timeline(None)
        dbg.svae_debug_set_b1_knock(0)
        return tl.cpu().double()

    def report(t, title):
        used = t[:, 0, 11] > 0
        t = t[used]
        print(f'--- {title}: shape [{B},{H},{L},64]: {int(used.sum())} CTAs, tiles per CTA {t[:, 0, 11].min():.0f}..{t[:, 0, 11].max():.0f}, '
              f'cycles per CTA mean {t[:, 0, 10].mean():.0f} max {t[:, 0, 10].max():.0f} -> {t[:, 0, 10].mean() / t[:, 0, 11].mean():.0f} cycles per tile')
        print('%-14s %9s | ' % ('role', 'busy') + ' '.join('%9s' % k for k in KINDS) + '   (cycles per tile, mean over CTAs)')
        for w, name in ROLES.items():
            per_tile = t[:, w, :10].mean(0) / t[:, 0, 11].mean()
            total = t[:, w, 10].mean() / t[:, 0, 11].mean()
            print('%-14s %9.0f | ' % (name, total - per_tile.sum()) + ' '.join('%9.0f' % x for x in per_tile.tolist()))
        for w in (0, 1, 3, 4, 7):
            x = t[:, w].mean(0)
            print(f'{ROLES[w]}: live units per tile {x[13] / x[11]:.2f}, cycles per live unit (ld + math + tmem st) {x[12] / max(x[13], 1):.0f}, '
                  f'dead-unit cycles per tile {x[14] / x[11]:.0f}, fence + wait::st + arrive per tile {x[15] / x[11]:.0f}')

    # knock-out experiments: event-timed kernel with one part of the work removed (results are wrong on purpose)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    names = {0: 'full kernel', 1: 'math: every unit dead', 2: 'no dQ / global-block MMAs', 4: 'no drain / stores', 8: 'no dV / dK MMAs',
             16: 'S^T / dP^T one k-step', 32: 'no O loads (statistics)', 1 | 2 | 8: 'no math, only S^T / dP^T MMAs', 2 | 8 | 16: 'math only',
             1 | 2 | 4 | 8 | 16 | 32: 'skeleton (loads + barriers)', 63 | 64: 'skeleton without TMA loads', 63 | 128: 'skeleton without math stores / fences',
             63 | 256: 'skeleton without any MMA', 63 | 64 | 128 | 256: 'barrier choreography only'}
    for mask, name in names.items():
        dbg.svae_debug_set_b1_knock(mask)
        ts = []
        for it in range(3):
            flush.zero_()
            ev0.record()
            rc = dbg.svae_attn_bwd(ctypes.byref(desc), q.data_ptr(), k.data_ptr(), v.data_ptr(), out2.data_ptr(), dout.data_ptr(),
                                   lse.data_ptr(), None, dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ws_ptr, ws_bytes, st)
            ev1.record()
            assert rc == 0, dbg.svae_last_error()
            torch.cuda.synchronize()
            ts.append(ev0.elapsed_time(ev1) * 1e3)
        print(f'knock {mask:3d} {name:38s}: {min(ts):7.1f} us (both kernels incl. finish)')
    dbg.svae_debug_set_b1_knock(0)
    if '--timeline' in sys.argv:
        report(with_timeline(0), 'full kernel')
        report(with_timeline(63), 'skeleton (knock 63)')


if __name__ == '__main__':
    main()
