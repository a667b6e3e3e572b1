"""Oracle (test infrastructure): block-sparse attention of `SparseAttention.__call__`.

Reference call site: sparse_vae/core/sparse_attention.py:75-92
    scores = sdd(q, k)                       # trans_b=True, one 32x32 tile per non-zero layout block
    dist   = softmax(scores, scale=Dh**-0.5, is_causal=self.causal,
                     key_padding_mask=kpm.half(), attn_mask=None)
    out    = dsd(dist, v)

The two ops are `triton.ops.blocksparse.{matmul,softmax}` from the pinned wheel triton==1.1.0
(requirements.txt:11), which is NOT under /root/reference (an edited copy of the matmul is vendored
at sparse_vae/core/sparse_matmul.py).  Published algorithm restated here:

  sdd      S[b, n] = Q[b, h_n, rows(r_n)] @ K[b, h_n, rows(c_n)]^T for the n-th non-zero block
           (h_n, r_n, c_n) in row-major order (sparse_matmul.py:19-92,133-144)
  softmax  per query row over that row's non-zero blocks, fp32:
           x = S*scale + key_padding_mask[b, key] (+ attn_mask); causal: key_pos > query_pos -> -inf;
           y = exp(x - max x) / sum
  dsd      O[b, h, rows(r)] = sum_{n in row r} P[b, n] @ V[b, h, rows(c_n)]   (sparse_matmul.py:152-219)

`blocksparse_attention` follows that structure literally (slow; small shapes).
`dense_masked_attention` is the algebraically identical dense form (excluded blocks contribute
exp(-inf)=0), differentiable, used for larger shapes and for gradients.
`attention_backward` restates the autograd of the three ops (sparse_matmul.py:463-488 plus the
softmax backward dx = y*(dy - sum(dy*y))*scale).
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import layout as _layout


def reference_kpm(bool_padding: torch.Tensor) -> torch.Tensor:
    """Additive key-padding mask exactly as the reference builds and casts it.

    core/attention.py:79 `mask * -1e7`, then sparse_attention.py:89 `.half()`: -1e7 overflows fp16
    to -inf, 0 stays 0.  Returned as fp32 so it can be added in fp32 math.
    """
    return (bool_padding * -1e7).half().float()


def expanded_mask(layout2d: np.ndarray, block: int, causal: bool, device=None) -> torch.Tensor:
    """[L, L] bool, True where attention is allowed (layout block non-zero and, if causal, key<=query)."""
    lay = torch.as_tensor(layout2d.astype(np.bool_), device=device)
    allowed = lay.repeat_interleave(block, 0).repeat_interleave(block, 1)
    if causal:
        L = allowed.shape[0]
        allowed = allowed & torch.ones(L, L, dtype=torch.bool, device=device).tril()
    return allowed


def dense_masked_attention(q, k, v, layout2d: np.ndarray, block: int = 32, causal: bool = True,
                           key_padding_mask: Optional[torch.Tensor] = None,
                           scale: Optional[float] = None, return_lse: bool = False):
    """q,k,v: [B,H,L,Dh] (any float dtype; math in that dtype, use fp32/fp64). kpm: additive [B,L]."""
    B, H, L, Dh = q.shape
    scale = Dh ** -0.5 if scale is None else scale
    scores = (q @ k.transpose(-1, -2)) * scale
    if key_padding_mask is not None:
        scores = scores + key_padding_mask[:, None, None, :].to(scores.dtype)
    allowed = expanded_mask(layout2d, block, causal, device=q.device)
    scores = scores.masked_fill(~allowed, float('-inf'))
    probs = scores.softmax(dim=-1)
    out = probs @ v
    if return_lse:
        return out, torch.logsumexp(scores, dim=-1)
    return out


def blocksparse_attention(q, k, v, layout2d: np.ndarray, block: int = 32, causal: bool = True,
                          key_padding_mask: Optional[torch.Tensor] = None, scale: Optional[float] = None):
    """Literal sdd -> softmax -> dsd over the non-zero blocks (same layout for every head)."""
    B, H, L, Dh = q.shape
    nb = L // block
    assert layout2d.shape == (nb, nb)
    scale = Dh ** -0.5 if scale is None else scale
    row_ptr, col_idx = _layout.csr(layout2d)
    out = torch.zeros_like(q)
    pos = torch.arange(block, device=q.device)
    for r in range(nb):
        cols = col_idx[row_ptr[r]:row_ptr[r + 1]]
        if len(cols) == 0:
            out[:, :, r * block:(r + 1) * block] = float('nan')   # softmax over an empty row
            continue
        qr = q[:, :, r * block:(r + 1) * block]                                  # [B,H,32,Dh]
        # sdd: one tile per non-zero block of this block-row
        tiles = []
        for c in cols:
            kc = k[:, :, c * block:(c + 1) * block]
            s = (qr @ kc.transpose(-1, -2)) * scale                             # [B,H,32,32]
            if key_padding_mask is not None:
                s = s + key_padding_mask[:, None, None, c * block:(c + 1) * block].to(s.dtype)
            if causal:
                qpos = r * block + pos[:, None]
                kpos = c * block + pos[None, :]
                s = s.masked_fill(kpos > qpos, float('-inf'))
            tiles.append(s)
        # softmax across the row's tiles
        x = torch.cat(tiles, dim=-1)                                             # [B,H,32,32*n]
        p = torch.softmax(x, dim=-1)
        # dsd
        acc = torch.zeros_like(qr)
        for j, c in enumerate(cols):
            vc = v[:, :, c * block:(c + 1) * block]
            acc = acc + p[..., j * block:(j + 1) * block] @ vc
        out[:, :, r * block:(r + 1) * block] = acc
    return out


def attention_backward(q, k, v, dout, layout2d: np.ndarray, block: int = 32, causal: bool = True,
                       key_padding_mask: Optional[torch.Tensor] = None, scale: Optional[float] = None):
    """Explicit gradients (dq, dk, dv): dV=P^T dO, dP=dO V^T, dS=P*(dP-rowsum(dP*P))*scale, dQ=dS K, dK=dS^T Q."""
    B, H, L, Dh = q.shape
    scale = Dh ** -0.5 if scale is None else scale
    scores = (q @ k.transpose(-1, -2)) * scale
    if key_padding_mask is not None:
        scores = scores + key_padding_mask[:, None, None, :].to(scores.dtype)
    allowed = expanded_mask(layout2d, block, causal, device=q.device)
    scores = scores.masked_fill(~allowed, float('-inf'))
    p = scores.softmax(dim=-1)
    dv = p.transpose(-1, -2) @ dout
    dp = dout @ v.transpose(-1, -2)
    ds = p * (dp - (dp * p).sum(dim=-1, keepdim=True)) * scale
    dq = ds @ k
    dk = ds.transpose(-1, -2) @ q
    return dq, dk, dv


def algorithmic_flops_fwd(B: int, H: int, nnz_per_head: int, block: int, Dh: int) -> int:
    """two block x block x Dh GEMMs per non-zero block."""
    return B * H * nnz_per_head * 4 * block * block * Dh


def algorithmic_bytes_fwd(B: int, H: int, L: int, Dh: int, elem_size: int = 2) -> int:
    """read Q,K,V and write O once."""
    return 4 * B * L * H * Dh * elem_size
