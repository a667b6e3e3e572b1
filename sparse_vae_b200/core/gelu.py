"""erf-GELU and the feed-forward block built on it (reference core/transformer_layer.py:20-24:
`nn.Sequential(nn.Linear(d, 4d), nn.GELU(), nn.Linear(4d, d))`; core/transformer_language_model.py:58 `output_layer`).

`GELU` is a drop-in `nn.GELU` (no parameters, same state_dict) whose 16-bit CUDA path is csrc/gelu.cu: HBM-bound forward
and backward (ATen's are ALU-bound on erff: 148 us / ~200 us at [65536, 2048] bf16 against 84 / 125 us of traffic).

`ffn_forward(seq, x)` runs a whole `Sequential(Linear, GELU, Linear)` as ONE autograd node under 16-bit autocast: the
four library GEMMs of `core/linear.py` (weight gradients in fp32, 16-bit weight shadows) around the GELU kernels, with
the up-projection's bias gradient taken from the column sums the GELU backward kernel accumulates on its way --
`svae_colsum` does not read the [rows, 4 d_model] gradient a second time -- and that gradient written in place over
the down-projection's input gradient.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn, Tensor

from .. import _native as N
from .linear import Linear, WeightShadows, colsum, _MIN_ROWS

_COUNTERS: dict = {}          # (device, stream) -> zeroed uint32 tickets (left at zero by every launch)


def _supported(x: Tensor) -> bool:
    return bool(N.FUSED_EXTRAS and x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and x.ndim >= 1
                and x.numel() > 0 and x.shape[-1] % 8 == 0)


def gelu_forward(x: Tensor) -> Tensor:
    """y = x * Phi(x) of a contiguous 16-bit CUDA tensor (csrc/gelu.cu)."""
    x = x.contiguous()
    y = torch.empty_like(x)
    n = x.shape[-1]
    N.check(N.lib.svae_gelu_fwd(x.data_ptr(), y.data_ptr(), N.svae_dtype(x.dtype), x.numel() // n, n,
                                N.current_stream(x.device)), 'svae_gelu_fwd')
    return y


def gelu_backward(dy: Tensor, x: Tensor, want_colsum: bool = False, inplace: bool = False):
    """(dx, column sums of dx | None): dx = dy * gelu'(x); `inplace` writes dx over dy."""
    dy = dy.contiguous()
    n = x.shape[-1]
    rows = x.numel() // n
    dx = dy if inplace else torch.empty_like(dy)
    sums = ws = counters = None
    ws_floats = 0
    stream = N.current_stream(x.device)
    if want_colsum:
        sums = torch.empty(n, device=x.device, dtype=torch.float32)
        ws_floats = N.lib.svae_gelu_bwd_workspace_floats(rows, n)
        ws = torch.empty(ws_floats, device=x.device, dtype=torch.float32)
        key = (x.device, stream)
        counters = _COUNTERS.get(key)
        need = N.lib.svae_gelu_bwd_counters(n)
        if counters is None or counters.numel() < need:
            counters = _COUNTERS[key] = torch.zeros(max(1024, need), device=x.device, dtype=torch.int32)
    N.check(N.lib.svae_gelu_bwd(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), N.svae_dtype(x.dtype), rows, n,
                                sums.data_ptr() if want_colsum else None, ws.data_ptr() if want_colsum else None, ws_floats,
                                counters.data_ptr() if want_colsum else None, stream), 'svae_gelu_bwd')
    return dx, sums


class _GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return gelu_forward(x)

    @staticmethod
    def backward(ctx, g: Tensor):
        (x,) = ctx.saved_tensors
        return gelu_backward(g, x)[0]


class GELU(nn.GELU):
    """`nn.GELU()` (erf form); 16-bit CUDA tensors go through csrc/gelu.cu."""

    def forward(self, x: Tensor) -> Tensor:
        if self.approximate == 'none' and _supported(x):
            return _GeluFn.apply(x)
        return super().forward(x)


class _FfnFn(torch.autograd.Function):
    """down(gelu(up(x))) under 16-bit autocast.  Saved: the 16-bit input, the up-projection's output and its GELU (what
    the three separate nodes of the reference save)."""

    @staticmethod
    def forward(ctx, x: Tensor, dtype: torch.dtype, shadows, w1, b1, w1_16, b1_16, w2, b2, w2_16, b2_16):
        x16 = x.to(dtype)
        if w1_16 is None:
            w1_16, b1_16 = w1.to(dtype), (b1.to(dtype) if b1 is not None else None)
            w2_16, b2_16 = w2.to(dtype), (b2.to(dtype) if b2 is not None else None)
        h = F.linear(x16, w1_16, b1_16)
        a = gelu_forward(h)
        ctx.save_for_backward(x16, h, a, w1_16, w2_16)
        ctx.in_dtypes = (x.dtype, w1.dtype, b1.dtype if b1 is not None else None, b2.dtype if b2 is not None else None)
        ctx.shadows, ctx.epoch = shadows, (shadows.epoch if shadows is not None else 0)
        return F.linear(a, w2_16, b2_16)

    @staticmethod
    def backward(ctx, g: Tensor):
        x16, h, a, w1_16, w2_16 = ctx.saved_tensors
        if ctx.shadows is not None and ctx.shadows.epoch != ctx.epoch:
            raise RuntimeError("a weight needed for this backward pass was modified (optimizer step?) after the forward "
                               "pass that used its 16-bit shadow copy")
        xd, wd, b1d, b2d = ctx.in_dtypes
        need = ctx.needs_input_grad
        d_out, d_hidden = w2_16.shape
        d_in = w1_16.shape[1]
        g2 = g.reshape(-1, d_out)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        dw2 = db2 = dw1 = db1 = dx = None
        if need[7]:
            dw2 = torch.mm(g2.t(), a.reshape(-1, d_hidden), out_dtype=torch.float32).to(wd)
        if b2d is not None and need[8]:
            db2 = colsum(g2).to(b2d)
        if need[0] or need[3] or (b1d is not None and need[4]):
            da = torch.mm(g2, w2_16)                                              # gradient of the GELU output
            want = b1d is not None and need[4]
            dh, sums = gelu_backward(da, h.reshape(-1, d_hidden), want_colsum=want, inplace=True)
            if want:
                db1 = sums.to(b1d)
            if need[3]:
                dw1 = torch.mm(dh.t(), x16.reshape(-1, d_in), out_dtype=torch.float32).to(wd)
            if need[0]:
                dx = torch.mm(dh, w1_16).view(x16.shape).to(xd)
        return dx, None, None, dw1, db1, None, None, dw2, db2, None, None


def ffn_forward(seq: nn.Sequential, x: Tensor) -> Tensor:
    """`seq(x)` for a `Sequential(Linear, GELU, Linear)`; one fused autograd node where csrc/gelu.cu applies."""
    if len(seq) == 3 and isinstance(seq[0], Linear) and isinstance(seq[2], Linear) and isinstance(seq[1], nn.GELU) \
            and seq[1].approximate == 'none' and seq[0]._fused_ok(x) and seq[0].out_features % 8 == 0 \
            and x.numel() >= _MIN_ROWS * seq[0].in_features:
        dtype = torch.get_autocast_dtype('cuda')
        if dtype in (torch.bfloat16, torch.float16):
            up, down = seq[0], seq[2]
            s1, s2 = up._active_shadow(dtype), down._active_shadow(dtype)
            use = s1 is not None and s2 is not None
            return _FfnFn.apply(x, dtype, s1[0] if use else None,
                                up.weight, up.bias, s1[1] if use else None, s1[2] if use else None,
                                down.weight, down.bias, s2[1] if use else None, s2[2] if use else None)
    return seq(x)
