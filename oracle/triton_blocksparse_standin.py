"""Oracle (test infrastructure): CPU stand-in for `triton.ops.blocksparse.{matmul, softmax}`.

The reference imports these from the pinned wheel triton==1.1.0 (requirements.txt:11,
sparse_vae/core/sparse_attention.py:3) which is absent from /root/reference and cannot run without a
GPU.  This module restates their published behaviour in differentiable torch ops with the same
constructor / call signatures the reference uses (sparse_attention.py:64,71,84-92), including the
block-sparse intermediate tensor format `[B, nnz_total, block, block]` with non-zero blocks enumerated
in row-major `(head, row, col)` order (vendored copy: sparse_vae/core/sparse_matmul.py:83-91,295-302).

It exists so that the reference's own, unmodified `SparseAttention.__call__` can be executed by
oracle/reference_harness.py when generating the golden fixtures.
"""
from __future__ import annotations

import torch


class matmul:
    def __init__(self, layout, block, mode, trans_a=False, trans_b=False):
        assert mode in ('sdd', 'dsd'), "only the modes the reference uses (sparse_attention.py:81,83)"
        self.layout = layout.clone()
        self.block = block
        self.mode = mode
        self.trans_a, self.trans_b = trans_a, trans_b
        idx = self.layout.nonzero()            # row-major (head, row, col)
        self.h, self.r, self.c = idx[:, 0], idx[:, 1], idx[:, 2]

    def __call__(self, a, b):
        blk = self.block
        if self.mode == 'sdd':
            assert self.trans_b and not self.trans_a
            B, H, L, Dh = a.shape
            nb = L // blk
            ab = a.reshape(B, H, nb, blk, Dh)
            bb = b.reshape(B, H, nb, blk, Dh)
            qa = ab[:, self.h, self.r]                      # [B, nnz, blk, Dh]
            kb = bb[:, self.h, self.c]
            return qa @ kb.transpose(-1, -2)                # [B, nnz, blk, blk]
        else:
            # a: sparse [B, nnz, blk, blk], b: dense [B, H, L, Dh]
            B, H, L, Dh = b.shape
            nb = L // blk
            vb = b.reshape(B, H, nb, blk, Dh)[:, self.h, self.c]     # [B, nnz, blk, Dh]
            prod = a @ vb                                            # [B, nnz, blk, Dh]
            out = torch.zeros(B, H * nb, blk, Dh, dtype=prod.dtype, device=prod.device)
            out = out.index_add(1, self.h * nb + self.r, prod)
            return out.reshape(B, H, nb * blk, Dh)


class softmax:
    def __init__(self, layout, block):
        self.layout = layout.clone()
        self.block = block
        idx = self.layout.nonzero()
        self.h, self.r, self.c = idx[:, 0], idx[:, 1], idx[:, 2]

    def __call__(self, x, scale=1.0, rpe=None, key_padding_mask=None, attn_mask=None,
                 key_padding_mask_mode='add', attn_mask_mode='add', is_causal=False):
        assert rpe is None and attn_mask is None, "never passed by the reference (core/attention.py:81)"
        assert key_padding_mask_mode == 'add'
        blk = self.block
        B, nnz = x.shape[:2]
        H, nb = self.layout.shape[0], self.layout.shape[1]
        xf = x.float() * scale
        if key_padding_mask is not None:
            kpm = key_padding_mask.float().reshape(B, nb, blk)[:, self.c]      # [B, nnz, blk]
            xf = xf + kpm[:, :, None, :]
        if is_causal:
            pos = torch.arange(blk, device=x.device)
            qpos = self.r[:, None, None] * blk + pos[None, :, None]
            kpos = self.c[:, None, None] * blk + pos[None, None, :]
            xf = xf.masked_fill((kpos > qpos)[None], float('-inf'))
        # scatter to dense rows, softmax, gather back
        dense = torch.full((B, H, nb, blk, nb, blk), float('-inf'), dtype=xf.dtype, device=x.device)
        dense[:, self.h, self.r, :, self.c, :] = xf.permute(1, 0, 2, 3)
        probs = dense.reshape(B, H, nb, blk, nb * blk).softmax(dim=-1).reshape(B, H, nb, blk, nb, blk)
        out = probs[:, self.h, self.r, :, self.c, :].permute(1, 0, 2, 3)
        return out.to(x.dtype)
