"""`LayerNorm`: drop-in subclass of `nn.LayerNorm` (same parameters / state_dict keys) backed by the one-pass
sm_100a kernels of csrc/layernorm.cu.  Used by `TransformerLayer` (reference core/transformer_layer.py:17-24:
`attn_layer_norm`, `ffn_layer_norm`, ...) and the output head (core/transformer_language_model.py:55-63).

Arithmetic is the reference's: fp32 statistics and affine transform.  Under autocast the reference's LayerNorm
returns fp32 and each consuming `nn.Linear` casts that to the autocast dtype; every LayerNorm of the model feeds
only Linear layers, so this module writes the rounded 16-bit result directly (`emit_autocast_dtype=True`) -- the
values the Linear layers see are bit-identical, the three per-consumer casts and the fp32 round trip disappear.
(The only numerical difference: gradients of several consumers are summed in the 16-bit dtype before the
LayerNorm backward instead of in fp32.)  Shapes / dtypes the kernels do not cover use ATen's `F.layer_norm`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn, Tensor

from .. import _native as N


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias, eps: float, out_dtype: torch.dtype):
        n = x.shape[-1]
        x2 = x.reshape(-1, n)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        rows = x2.shape[0]
        y = torch.empty((rows, n), device=x.device, dtype=out_dtype)
        stats = torch.empty((2, rows), device=x.device, dtype=torch.float32)
        N.check(N.lib.svae_layernorm_fwd(x2.data_ptr(), N.svae_dtype(x2.dtype), weight.data_ptr(), N.ptr(bias), rows, n,
                                         float(eps), y.data_ptr(), N.svae_dtype(out_dtype), stats[0].data_ptr(),
                                         stats[1].data_ptr(), N.current_stream(x.device)), 'svae_layernorm_fwd')
        ctx.save_for_backward(x2, weight, stats)
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy: Tensor):
        dx, dgamma, dbeta, _ = _layer_norm_backward(ctx, ctx.saved_tensors, dy, None)
        return dx, dgamma, dbeta, None, None


def _layer_norm_backward(ctx, saved, dy: Tensor, dres, low_dtype=None):
    """One pass over x and dy: dx (+ dres, the gradient arriving around the norm), dgamma, dbeta; with `low_dtype`
    also dx rounded to that 16-bit dtype (fourth result).
    `saved` = ctx.saved_tensors, read exactly once by the caller (activation checkpointing insists on that)."""
    x2, weight, stats = saved
    rows, n = x2.shape
    dy2 = dy.reshape(rows, n)
    if not dy2.is_contiguous():
        dy2 = dy2.contiguous()
    need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
    want_low = low_dtype is not None and low_dtype == dy2.dtype
    dx = torch.empty_like(x2) if (need_x or want_low) else None
    dx_low = torch.empty(x2.shape, device=x2.device, dtype=low_dtype) if want_low else None
    dgb = torch.empty((2, n), device=x2.device, dtype=torch.float32) if (need_w or need_b) else None
    ws_floats = N.lib.svae_layernorm_bwd_workspace_floats(rows, n)
    ws = torch.empty(ws_floats, device=x2.device, dtype=torch.float32)
    N.check(N.lib.svae_layernorm_bwd(dy2.data_ptr(), N.svae_dtype(dy2.dtype), x2.data_ptr(), N.svae_dtype(x2.dtype),
                                     weight.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), rows, n, N.ptr(dx),
                                     N.ptr(dres) if dx is not None else None, N.ptr(dx_low),
                                     dgb[0].data_ptr() if need_w else None, dgb[1].data_ptr() if need_b else None,
                                     ws.data_ptr(), ws_floats, N.current_stream(x2.device)), 'svae_layernorm_bwd')
    return (dx.view(ctx.x_shape) if dx is not None else None, dgb[0] if need_w else None, dgb[1] if need_b else None,
            dx_low.view(ctx.x_shape) if dx_low is not None else None)


class _NormForkFn(torch.autograd.Function):
    """(x, LayerNorm(x)) for the pre-norm residual blocks (reference core/transformer_layer.py:35-61: x feeds both the
    norm of a sub-layer and the residual connection around it).  Backward adds the two gradients of x inside the
    LayerNorm-backward pass instead of leaving the sum to a separate accumulation kernel."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias, eps: float, out_dtype: torch.dtype):
        y = _LayerNormFn.forward(ctx, x, weight, bias, eps, out_dtype)
        return x.view_as(x), y

    @staticmethod
    def backward(ctx, g_skip, dy):
        if dy is None:
            return g_skip, None, None, None, None
        saved = ctx.saved_tensors
        dres = None
        if g_skip is not None and ctx.needs_input_grad[0]:
            x2 = saved[0]
            dres = g_skip.reshape(x2.shape)
            if dres.dtype != x2.dtype or not dres.is_contiguous():
                dres = dres.to(x2.dtype).contiguous()
        dx, dgamma, dbeta, _ = _layer_norm_backward(ctx, saved, dy, dres)
        return dx, dgamma, dbeta, None, None


class _AddNormFn(torch.autograd.Function):
    """(x + h, LayerNorm(x + h)) for the fp32 stream x and a 16-bit branch output h: the residual update and the next
    sub-layer's norm in one launch (reference core/transformer_layer.py:41-61).  Backward: one LayerNorm-backward
    pass yields the stream gradient (skip-path gradient added inside) and its 16-bit copy for the branch."""

    @staticmethod
    def forward(ctx, x: Tensor, h: Tensor, weight: Tensor, bias, eps: float):
        n = x.shape[-1]
        rows = x.numel() // n
        x_new = torch.empty_like(x)
        y = torch.empty_like(h)
        stats = torch.empty((2, rows), device=x.device, dtype=torch.float32)
        N.check(N.lib.svae_residual_layernorm(x.data_ptr(), h.data_ptr(), N.svae_dtype(h.dtype), weight.data_ptr(), N.ptr(bias),
                                              rows, n, float(eps), y.data_ptr(), N.svae_dtype(h.dtype), x_new.data_ptr(),
                                              stats[0].data_ptr(), stats[1].data_ptr(), N.current_stream(x.device)),
                'svae_residual_layernorm')
        ctx.save_for_backward(x_new.view(rows, n), weight, stats)
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        ctx.h_dtype = h.dtype
        return x_new, y

    @staticmethod
    def backward(ctx, g_skip, dy):
        need_x, need_h = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if dy is None:                                   # the norm's output was not used
            return (g_skip if need_x else None, g_skip.to(ctx.h_dtype) if need_h and g_skip is not None else None,
                    None, None, None)
        saved = ctx.saved_tensors
        dres = None
        if g_skip is not None:
            dres = g_skip.reshape(saved[0].shape)
            if dres.dtype != torch.float32 or not dres.is_contiguous():
                dres = dres.float().contiguous()
        # needs_input_grad is indexed (x, h, weight, bias, eps) here; the shared helper expects (x, weight, bias)
        ctx_view = _Needs(need_x or need_h, ctx.needs_input_grad[2], ctx.needs_input_grad[3], ctx.has_bias, ctx.x_shape)
        dx, dgamma, dbeta, dx_low = _layer_norm_backward(ctx_view, saved, dy, dres, ctx.h_dtype if need_h else None)
        if need_h and dx_low is None:                    # dy arrived in another dtype than the branch
            dx_low = dx.to(ctx.h_dtype)
        return dx if need_x else None, dx_low, dgamma, dbeta, None


class _Needs:
    """The fields `_layer_norm_backward` reads from an autograd ctx, for a Function with a different argument list."""

    def __init__(self, need_x, need_w, need_b, has_bias, x_shape):
        self.needs_input_grad = (need_x, need_w, need_b)
        self.has_bias = has_bias
        self.x_shape = x_shape


_PAIRS = {(torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.float32, torch.float16),
          (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32), (torch.float16, torch.float16),
          (torch.float16, torch.float32)}


class LayerNorm(nn.LayerNorm):
    emit_autocast_dtype: bool = True

    def _fused_out_dtype(self, x: Tensor):
        fused = (N.FUSED_EXTRAS and x.is_cuda and self.elementwise_affine and len(self.normalized_shape) == 1 and x.numel() > 0
                 and self.weight.dtype == torch.float32 and N.lib.svae_layernorm_supported(x.shape[-1]))
        if fused:
            out_dtype = x.dtype
            if torch.is_autocast_enabled('cuda'):
                out_dtype = torch.get_autocast_dtype('cuda') if self.emit_autocast_dtype else torch.float32
            if (x.dtype, out_dtype) in _PAIRS:
                return out_dtype
        return None

    def forward(self, x: Tensor) -> Tensor:
        out_dtype = self._fused_out_dtype(x)
        if out_dtype is not None:
            return _LayerNormFn.apply(x, self.weight, self.bias, self.eps, out_dtype)
        return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)

    def add_fork(self, x: Tensor, h: Tensor):
        """(x + h, self(x + h)): residual update by a 16-bit branch output and this norm in one launch; the returned
        stream carries the skip path like `fork`."""
        if (self._fused_out_dtype(x) == h.dtype and x.dtype == torch.float32 and h.dtype in (torch.bfloat16, torch.float16)
                and x.shape == h.shape and torch.is_grad_enabled() and (x.requires_grad or h.requires_grad)
                and x.is_contiguous() and h.is_contiguous()):
            return _AddNormFn.apply(x, h, self.weight, self.bias, self.eps)
        from .residual import residual_add
        return self.fork(residual_add(x, h))

    def fork(self, x: Tensor):
        """(x, self(x)) for a pre-norm block whose residual connection goes around this norm: use the returned x for
        the skip path, so that backward sums the two gradients of x inside the LayerNorm-backward kernel."""
        out_dtype = self._fused_out_dtype(x)
        if out_dtype is not None and torch.is_grad_enabled() and x.requires_grad:
            return _NormForkFn.apply(x, self.weight, self.bias, self.eps, out_dtype)
        return x, self(x)
