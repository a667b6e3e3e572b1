"""GPU parity of the few-queries-over-long-keys attention (csrc/xattn_sm100.cu, the Perceiver encoder's learned-query
layers) against float64 torch on the same (rounded) inputs; tolerance 1e-2 relative like the sparse kernels in bf16."""
import sys
from pathlib import Path

import pytest
import torch

sys.path.insert(0, str(Path(__file__).parent))
from util import rel_err  # noqa: E402

pytestmark = pytest.mark.gpu


def _ref(q, k, v, kpm, dout):
    qd, kd, vd = (t.detach().double().cpu().requires_grad_(True) for t in (q, k, v))
    s = qd @ kd.transpose(-1, -2) * q.shape[-1] ** -0.5
    if kpm is not None:
        s = s + kpm.double().cpu()[:, None, None, :]
    out = s.softmax(-1) @ vd
    out.backward(dout.double().cpu())
    return out.detach(), qd.grad, kd.grad, vd.grad


@pytest.mark.parametrize('B,H,nq,Lk,dtype,lengths', [
    (2, 8, 64, 512, torch.bfloat16, None), (2, 8, 64, 4096, torch.bfloat16, [4096, 3001]), (1, 2, 64, 300, torch.bfloat16, None),
    (3, 4, 40, 1000, torch.bfloat16, [1000, 77, 640]), (1, 8, 1, 384, torch.bfloat16, None), (2, 8, 64, 1024, torch.float16, [1024, 900]),
    (16, 8, 64, 4096, torch.bfloat16, None),
])
def test_cross_attention_matches_float64(B, H, nq, Lk, dtype, lengths):
    from sparse_vae_b200.core import cross_attention as xa
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(nq + Lk)
    # the encoder's layout: k / v are [B, H, Lk, 64] views of [B, Lk, H*64]; q the expanded learned queries
    kk, vv = (torch.randn(B, Lk, H * 64, generator=g).to(dev, dtype) for _ in range(2))
    k, v = (t.unflatten(-1, (H, 64)).transpose(1, 2).requires_grad_(True) for t in (kk, vv))
    q = (torch.randn(1, nq, H * 64, generator=g) * 2).to(dev, dtype).expand(B, nq, H * 64).unflatten(-1, (H, 64)).transpose(1, 2)
    q = q.detach().requires_grad_(True)
    dout = torch.randn(B, nq, H * 64, generator=g).to(dev, dtype).unflatten(-1, (H, 64)).transpose(1, 2)
    kpm = None
    if lengths:
        pad = torch.zeros(B, Lk, dtype=torch.bool)
        for b, n in enumerate(lengths):
            pad[b, n:] = True
        kpm = (pad * -1e7).to(dev)
    assert xa.supported(q, k, v) == (B * H >= xa.MIN_PAIRS)
    out = xa.cross_attention(q, k, v, kpm)
    assert out.shape == q.shape and out.dtype == dtype
    out.backward(dout)
    small = B * H * Lk <= 2 * 8 * 4096
    sl = (slice(None),) if small else (slice(B - 1, B), slice(3, 4))           # the float64 reference of a slice is enough at full size
    ro, rdq, rdk, rdv = _ref(q[sl], k[sl], v[sl], None if kpm is None else kpm[sl[0]], dout[sl])
    errs = dict(out=rel_err(out[sl], ro), dq=rel_err(q.grad[sl], rdq), dk=rel_err(k.grad[sl], rdk), dv=rel_err(v.grad[sl], rdv))
    assert all(e <= 1e-2 for e in errs.values()), errs
    # deterministic (no atomics)
    q.grad = k.grad = v.grad = None
    out2 = xa.cross_attention(q, k, v, kpm)
    out2.backward(dout)
    assert torch.equal(out, out2)


def test_encoder_uses_the_kernel_and_matches_sdpa():
    """Perceiver first layer through Attention.forward: same numbers as the library SDPA path it replaces."""
    import sparse_vae_b200 as sv
    from sparse_vae_b200 import _native as N
    from sparse_vae_b200.core.lightning_shim import to_attrdict
    dev = torch.device('cuda')
    torch.manual_seed(3)
    model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams(num_layers=4))).to(dev).eval()
    model.initialize_weights()
    x = torch.randn(8, 1024, 512, device=dev)               # 8 x 8 heads = 64 (batch, head) pairs: cross_attention.MIN_PAIRS
    pad = torch.zeros(8, 1024, dtype=torch.bool, device=dev)
    pad[1, 800:] = True
    pad[5, 300:] = True
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
        N.profile_begin()
        a = model.encoder(x, padding=pad)
        prof = N.profile_end()
        assert prof.get('xattn_fwd_sm100', {}).get('launches', 0) >= 1
        N.FUSED_EXTRAS = False
        try:
            b = model.encoder(x, padding=pad)
        finally:
            N.FUSED_EXTRAS = True
    assert rel_err(a.float(), b.float()) <= 2e-2
