"""`x + h` for the fp32 residual stream and a 16-bit branch output (reference core/transformer_layer.py:41,49,61).

Under autocast the transformer blocks add the bf16 / fp16 result of attention and feed-forward to the fp32 stream
that started at the embedding.  ATen evaluates that mixed-dtype addition with a non-vectorised kernel; csrc/residual.cu
is the same fp32 addition with 16-byte accesses.  Gradients are what autograd gives the plain expression: g for x,
g cast to the branch dtype for h."""
from __future__ import annotations

import torch
from torch import Tensor

from .. import _native as N


class _ResidualAddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, h: Tensor):
        out = torch.empty_like(x)
        N.check(N.lib.svae_residual_add(x.data_ptr(), h.data_ptr(), N.svae_dtype(h.dtype), out.data_ptr(), x.numel(),
                                        N.current_stream(x.device)), 'svae_residual_add')
        ctx.h_dtype = h.dtype
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        return (g if ctx.needs_input_grad[0] else None), (g.to(ctx.h_dtype) if ctx.needs_input_grad[1] else None)


def residual_add(x: Tensor, h: Tensor) -> Tensor:
    if (N.FUSED_EXTRAS and x.is_cuda and x.dtype == torch.float32 and h.dtype in (torch.bfloat16, torch.float16)
            and x.shape == h.shape and x.numel() > 0 and x.numel() % 8 == 0 and x.is_contiguous() and h.is_contiguous()
            and x.data_ptr() % 16 == 0 and h.data_ptr() % 16 == 0):
        return _ResidualAddFn.apply(x, h)
    return x + h
