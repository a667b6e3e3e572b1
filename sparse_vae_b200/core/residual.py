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


class _ResidualDropoutAddFn(torch.autograd.Function):
    """x + dropout(h) in one launch; the keep mask is regenerated from (seed, offset) in backward, never stored."""

    @staticmethod
    def forward(ctx, x: Tensor, h: Tensor, p: float, seed: int, offset: int, philox_dev: int):
        out = torch.empty_like(x)
        N.check(N.lib.svae_residual_dropout_add_g(x.data_ptr(), h.data_ptr(), N.svae_dtype(h.dtype), out.data_ptr(), x.numel(),
                                                  p, seed, offset, philox_dev, N.current_stream(x.device)), 'svae_residual_dropout_add')
        ctx.args = (p, seed, offset, philox_dev, h.dtype)
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        p, seed, offset, philox_dev, h_dtype = ctx.args
        dh = None
        if ctx.needs_input_grad[1]:
            g32 = g.contiguous()
            dh = torch.empty(g32.shape, dtype=h_dtype, device=g32.device)
            N.check(N.lib.svae_dropout_branch_grad_g(g32.data_ptr(), dh.data_ptr(), N.svae_dtype(h_dtype), g32.numel(), p, seed,
                                                     offset, philox_dev, N.current_stream(g32.device)), 'svae_dropout_branch_grad')
        return (g if ctx.needs_input_grad[0] else None), dh, None, None, None, None


def residual_dropout_add(x: Tensor, h: Tensor, dropout: torch.nn.Dropout) -> Tensor:
    """`x + dropout(h)` (reference core/transformer_layer.py:61).  Training mode on the GPU: one launch, Philox mask from
    the default CUDA generator's (seed, offset), which is advanced like any other random op; otherwise the plain
    expression (through `residual_add`)."""
    if (dropout.training and 0.0 < dropout.p < 1.0 and not dropout.inplace and N.FUSED_EXTRAS and x.is_cuda
            and x.dtype == torch.float32 and h.dtype in (torch.bfloat16, torch.float16) and x.shape == h.shape
            and x.numel() > 0 and x.numel() % 8 == 0 and x.is_contiguous() and h.is_contiguous()
            and x.data_ptr() % 16 == 0 and h.data_ptr() % 16 == 0):
        from .graph_step import StepPhilox
        if StepPhilox.active is not None:          # a step is being captured: {seed, base offset} are read on the device
            philox_dev, delta = StepPhilox.active.reserve(4)
            return _ResidualDropoutAddFn.apply(x, h, float(dropout.p), 0, delta, philox_dev)
        if not torch.cuda.is_current_stream_capturing():
            gen = torch.cuda.default_generators[x.device.index if x.device.index is not None else torch.cuda.current_device()]
            seed, offset = gen.initial_seed(), gen.get_offset()
            gen.set_offset(offset + 4)
            return _ResidualDropoutAddFn.apply(x, h, float(dropout.p), seed, offset, 0)
    return residual_add(x, dropout(h))
