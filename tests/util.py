import numpy as np
import torch

from oracle import attention as oattn
from oracle import layout as olayout


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| relative to max |b| (the scale-aware bound the parity targets are stated in)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    # rows whose every key is masked are NaN in the reference softmax (exp(-inf - -inf)); they must be NaN in
    # both tensors at the same places and are then left out of the norm
    both_nan = torch.isnan(a) & torch.isnan(b)
    if not torch.equal(torch.isnan(a), torch.isnan(b)):
        return float('inf')
    a, b = a.masked_fill(both_nan, 0.0), b.masked_fill(both_nan, 0.0)
    denom = b.abs().max().item()
    return ((a - b).abs().max().item()) / (denom if denom > 0 else 1.0)


def block_rel_err(a: torch.Tensor, b: torch.Tensor, block: int = 32, floor: float = 1e-3) -> float:
    """Worst per-block error of a [B, H, L, Dh] tensor: max|a-b| over each (batch, head, 32-row block) relative to
    max|b| over THAT block (not the whole tensor), so that a wrong low-magnitude region -- a mis-masked padded block,
    a tile edge -- cannot hide behind the tensor's largest entries.  Blocks whose reference is (nearly) zero are held
    to `floor` times the tensor's scale instead."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if not torch.equal(torch.isnan(a), torch.isnan(b)):
        return float('inf')
    nan = torch.isnan(b)
    a, b = a.masked_fill(nan, 0.0), b.masked_fill(nan, 0.0)
    B, H, L, D = b.shape
    nb = L // block
    d = (a - b).abs()[:, :, :nb * block].reshape(B, H, nb, block * D).amax(-1)
    scale = b.abs()[:, :, :nb * block].reshape(B, H, nb, block * D).amax(-1)
    gmax = b.abs().max().item()
    return (d / scale.clamp_min(floor * (gmax if gmax > 0 else 1.0))).max().item()


def make_qkv(B, H, L, Dh, dtype, device, seed=0, strided=True, requires_grad=False):
    """q, k, v as the reference hands them to the op: [B,H,L,Dh] views of [B,L,H*Dh] (core/attention.py:76)."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    out = []
    for _ in range(3):
        t = torch.randn(B, L, H * Dh, generator=g, dtype=torch.float32).to(device=device, dtype=dtype)
        t = t.unflatten(-1, (H, Dh)).transpose(1, 2)
        if not strided:
            t = t.contiguous()
        out.append(t.requires_grad_(requires_grad))
    return out


def oracle_attention(q, k, v, cfg, pad=None, dout=None):
    """fp64 CPU oracle on the (possibly 16-bit rounded) inputs; returns out (and dq, dk, dv when dout is given)."""
    B, H, L, Dh = q.shape
    lay = olayout.layout_2d(L // cfg.block_size, cfg.window_size, cfg.causal, cfg.include_cls)
    qd, kd, vd = (t.detach().double().cpu().requires_grad_(dout is not None) for t in (q, k, v))
    kpm = oattn.reference_kpm(pad.cpu()).double() if pad is not None else None
    out = oattn.dense_masked_attention(qd, kd, vd, lay, cfg.block_size, cfg.causal, kpm)
    if dout is None:
        return out
    out.backward(dout.detach().double().cpu())
    return out.detach(), qd.grad, kd.grad, vd.grad


def make_padding(B, L, lengths, device):
    pad = torch.zeros(B, L, dtype=torch.bool)
    for b, n in enumerate(lengths):
        pad[b, n:] = True
    return pad.to(device)
