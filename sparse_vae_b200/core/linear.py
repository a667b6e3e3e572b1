"""`Linear`: drop-in subclass of `nn.Linear` (same parameters / state_dict keys) for the projections of the decoder
blocks (reference core/attention.py:33-39 `q/k/v/output_linear`, core/transformer_layer.py:20-24 `ffn`).

Forward is the library GEMM the reference uses (`F.linear` on the autocast-dtype operands).  Backward under autocast:
the two library GEMMs, with the weight gradient accumulated and written in fp32 (the reference rounds it to the
autocast dtype before casting back), and the bias gradient from the two-stage column-sum kernel `svae_colsum`
(csrc/colsum.cu) instead of ATen's generic reduction (~1.6 TB/s at [65536, 512]).  Outside autocast, on the CPU, for
small inputs or without a bias it is exactly `nn.Linear`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn, Tensor

from .. import _native as N

_MIN_ROWS = 1024


def colsum(x2d: Tensor) -> Tensor:
    """fp32 column sums of a [rows, n] CUDA matrix (unit inner stride, n % 8 == 0)."""
    rows, n = x2d.shape
    out = torch.empty(n, device=x2d.device, dtype=torch.float32)
    ws_floats = N.lib.svae_colsum_workspace_floats(rows, n)
    ws = torch.empty(ws_floats, device=x2d.device, dtype=torch.float32)
    N.check(N.lib.svae_colsum(x2d.data_ptr(), N.svae_dtype(x2d.dtype), rows, n, x2d.stride(0), out.data_ptr(),
                              ws.data_ptr(), ws_floats, N.current_stream(x2d.device)), 'svae_colsum')
    return out


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor, dtype: torch.dtype):
        x16, w16 = x.to(dtype), weight.to(dtype)
        ctx.save_for_backward(x16, w16)
        ctx.in_dtypes = (x.dtype, weight.dtype, bias.dtype)
        return F.linear(x16, w16, bias.to(dtype))

    @staticmethod
    def backward(ctx, g: Tensor):
        x16, w16 = ctx.saved_tensors
        xd, wd, bd = ctx.in_dtypes
        n_out, n_in = w16.shape
        g2 = g.reshape(-1, n_out)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.mm(g2, w16).view(x16.shape).to(xd)
        if ctx.needs_input_grad[1]:
            dw = torch.mm(g2.t(), x16.reshape(-1, n_in), out_dtype=torch.float32).to(wd)
        if ctx.needs_input_grad[2]:
            db = colsum(g2).to(bd)
        return dx, dw, db, None


class Linear(nn.Linear):
    def forward(self, x: Tensor) -> Tensor:
        if (N.FUSED_EXTRAS and x.is_cuda and self.bias is not None and torch.is_autocast_enabled('cuda') and torch.is_grad_enabled()
                and self.out_features % 8 == 0 and x.numel() >= _MIN_ROWS * self.in_features
                and (x.requires_grad or self.weight.requires_grad)):
            dtype = torch.get_autocast_dtype('cuda')
            if dtype in (torch.bfloat16, torch.float16):
                return _LinearFn.apply(x, self.weight, self.bias, dtype)
        return F.linear(x, self.weight, self.bias)
