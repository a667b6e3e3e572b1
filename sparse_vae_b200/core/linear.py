"""`Linear`: drop-in subclass of `nn.Linear` (same parameters / state_dict keys) for the projections of the decoder
blocks (reference core/attention.py:33-39 `q/k/v/output_linear`, core/transformer_layer.py:20-24 `ffn`).

Forward is the library GEMM the reference uses (`F.linear` on the autocast-dtype operands).  Backward under autocast:
the two library GEMMs, with the weight gradient accumulated and written in fp32 (the reference rounds it to the
autocast dtype before casting back), and the bias gradient from the two-stage column-sum kernel `svae_colsum`
(csrc/colsum.cu) instead of ATen's generic reduction (~1.6 TB/s at [65536, 512]).  Outside autocast, on the CPU, for
small inputs it is exactly `nn.Linear`.

`WeightShadows` removes the per-weight autocast casts: under autocast every `nn.Linear` casts its fp32 weight (and
bias) to the 16-bit dtype once per step -- ~150 launches of a few microseconds each for this model.  The shadow set
holds one flat 16-bit buffer with a view per `Linear` parameter and refreshes ALL of them with a handful of
multi-tensor launches (`svae_multi_tensor_cast`, same rounding as `.to(dtype)`) when a step begins; while the
step's forward runs, `Linear` reads those views instead of casting.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import List, Optional

import torch
import torch.nn.functional as F
from torch import nn, Tensor

from .. import _native as N
from ..fused_optim import _PtrList

_MIN_ROWS = 1024


_COLSUM_COUNTERS: dict = {}          # (device, stream) -> zeroed uint32 tickets for the single-launch column sum


def colsum(x2d: Tensor) -> Tensor:
    """fp32 column sums of a [rows, n] CUDA matrix (unit inner stride, n % 8 == 0), one launch."""
    rows, n = x2d.shape
    out = torch.empty(n, device=x2d.device, dtype=torch.float32)
    ws_floats = N.lib.svae_colsum_workspace_floats(rows, n)
    ws = torch.empty(ws_floats, device=x2d.device, dtype=torch.float32)
    stream = N.current_stream(x2d.device)
    # the kernel leaves its tickets at zero, so one array serves every call issued on the same stream
    key = (x2d.device, stream)
    counters = _COLSUM_COUNTERS.get(key)
    if counters is None or counters.numel() < N.lib.svae_colsum_counters(n):
        counters = _COLSUM_COUNTERS[key] = torch.zeros(max(1024, N.lib.svae_colsum_counters(n)), device=x2d.device,
                                                       dtype=torch.int32)
    N.check(N.lib.svae_colsum(x2d.data_ptr(), N.svae_dtype(x2d.dtype), rows, n, x2d.stride(0), out.data_ptr(),
                              ws.data_ptr(), ws_floats, counters.data_ptr(), stream), 'svae_colsum')
    return out


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], dtype: torch.dtype, w16: Optional[Tensor],
                b16: Optional[Tensor], shadows: Optional['WeightShadows']):
        x16 = x.to(dtype)
        if w16 is None:
            w16 = weight.to(dtype)
            b16 = bias.to(dtype) if bias is not None else None
        ctx.save_for_backward(x16, w16)
        ctx.in_dtypes = (x.dtype, weight.dtype, bias.dtype if bias is not None else None)
        ctx.shadows, ctx.epoch = shadows, (shadows.epoch if shadows is not None else 0)
        return F.linear(x16, w16, b16)

    @staticmethod
    def backward(ctx, g: Tensor):
        x16, w16 = ctx.saved_tensors
        if ctx.shadows is not None and ctx.shadows.epoch != ctx.epoch:
            raise RuntimeError("a weight needed for this backward pass was modified (optimizer step?) after the forward "
                               "pass that used its 16-bit shadow copy")
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
        if bd is not None and ctx.needs_input_grad[2]:
            db = colsum(g2).to(bd)
        return dx, dw, db, None, None, None, None


class _Linear3Fn(torch.autograd.Function):
    """q / k / v projections of one input (self-attention): three library GEMMs forward; backward accumulates the
    three input-gradient products in the GEMM epilogue (beta = 1) instead of two extra element-wise adds."""

    @staticmethod
    def forward(ctx, x: Tensor, dtype: torch.dtype, shadows, *wb):
        # wb = (w0, b0, w16_0, b16_0, w1, b1, w16_1, b16_1, w2, b2, w16_2, b16_2)
        x16 = x.to(dtype)
        outs, w16s = [], []
        for i in range(3):
            w, b, w16, b16 = wb[4 * i:4 * i + 4]
            if w16 is None:
                w16, b16 = w.to(dtype), b.to(dtype)
            w16s.append(w16)
            outs.append(F.linear(x16, w16, b16))
        ctx.save_for_backward(x16, *w16s)
        ctx.in_dtypes = (x.dtype, wb[0].dtype, wb[1].dtype)
        ctx.shadows, ctx.epoch = shadows, (shadows.epoch if shadows is not None else 0)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        x16, *w16s = ctx.saved_tensors
        if ctx.shadows is not None and ctx.shadows.epoch != ctx.epoch:
            raise RuntimeError("a weight needed for this backward pass was modified (optimizer step?) after the forward "
                               "pass that used its 16-bit shadow copy")
        xd, wd, bd = ctx.in_dtypes
        x2 = x16.reshape(-1, x16.shape[-1])
        dx = None
        grads = []
        for i, (g, w16) in enumerate(zip(gs, w16s)):
            dw = db = None
            if g is not None:
                g2 = g.reshape(-1, w16.shape[0])
                if not g2.is_contiguous():
                    g2 = g2.contiguous()
                if ctx.needs_input_grad[0]:
                    if dx is None:
                        dx = torch.mm(g2, w16)
                    else:
                        dx.addmm_(g2, w16)
                if ctx.needs_input_grad[3 + 4 * i]:
                    dw = torch.mm(g2.t(), x2, out_dtype=torch.float32).to(wd)
                if ctx.needs_input_grad[4 + 4 * i]:
                    db = colsum(g2).to(bd)
            grads += [dw, db, None, None]
        return (dx.view(x16.shape).to(xd) if dx is not None else None, None, None, *grads)


def _row_stride(t: Tensor):
    """Row stride (elements) if `t` [..., L, d] is a uniformly strided stack of rows with unit inner stride, else None."""
    if t.stride(-1) != 1 or t.ndim < 2:
        return None
    ld = t.stride(-2)
    for i in range(t.ndim - 2):
        if t.shape[i] > 1 and t.stride(i) != t.stride(i + 1) * t.shape[i + 1]:
            return None
    return ld if ld >= t.shape[-1] and ld % 8 == 0 and t.data_ptr() % 16 == 0 else None


def rotary_pair(a: Tensor, b: Tensor, cos: Tensor, sin: Tensor, conj: bool = False, want_colsum: bool = False,
                inplace: bool = False):
    """Both tensors ([..., L, d], same shape and row stride -- e.g. column slices of one [..., L, 3 d] buffer) rotated by
    one launch of `svae_rotary_pair`; with `want_colsum` also the fp32 column sums of the two results (bit-identical to
    `colsum` of them).  `inplace` rotates the rows where they are; otherwise the results are new contiguous tensors."""
    ld_a, ld_b = _row_stride(a), _row_stride(b)
    if ld_a is None or ld_a != ld_b:
        assert not inplace, "in-place rotation needs uniformly strided rows"
        a, b = a.contiguous(), b.contiguous()
        ld_a = a.shape[-1]
    L, d = a.shape[-2], a.shape[-1]
    rows = a.numel() // d
    if inplace:
        oa, ob, out_ld = a, b, ld_a
    else:
        oa, ob, out_ld = torch.empty(a.shape, dtype=a.dtype, device=a.device), torch.empty(b.shape, dtype=b.dtype, device=b.device), d
    stream = N.current_stream(a.device)
    sa = sb = ws = counters = None
    ws_floats = 0
    if want_colsum:
        sa = torch.empty(d, device=a.device, dtype=torch.float32)
        sb = torch.empty(d, device=a.device, dtype=torch.float32)
        ws_floats = N.lib.svae_rotary_pair_workspace_floats(rows, d)
        ws = torch.empty(ws_floats, device=a.device, dtype=torch.float32)
        key = (a.device, stream)
        counters = _COLSUM_COUNTERS.get(key)
        if counters is None or counters.numel() < N.lib.svae_colsum_counters(d):
            counters = _COLSUM_COUNTERS[key] = torch.zeros(max(1024, N.lib.svae_colsum_counters(d)), device=a.device,
                                                           dtype=torch.int32)
    N.check(N.lib.svae_rotary_pair(a.data_ptr(), b.data_ptr(), cos.data_ptr(), sin.data_ptr(), oa.data_ptr(), ob.data_ptr(),
                                   N.svae_dtype(a.dtype), N.svae_dtype(cos.dtype), rows, L, d, int(conj), ld_a, out_ld,
                                   sa.data_ptr() if want_colsum else None, sb.data_ptr() if want_colsum else None,
                                   ws.data_ptr() if want_colsum else None, ws_floats,
                                   counters.data_ptr() if want_colsum else None, stream), 'svae_rotary_pair')
    return oa, ob, sa, sb


def _adjacent(ts, numel_each: int) -> bool:
    """Three contiguous 16-bit tensors that sit back to back in one allocation."""
    return all(t is not None and t.is_contiguous() and t.numel() == numel_each for t in ts) and \
        all(ts[i + 1].data_ptr() == ts[i].data_ptr() + numel_each * ts[i].element_size() for i in range(2)) and \
        ts[0].untyped_storage().data_ptr() == ts[2].untyped_storage().data_ptr()


class _QkvRotaryFn(torch.autograd.Function):
    """The q / k / v projections of self-attention AND the rotary encoding of q and k as one autograd node (SURVEY 8f
    row 1; reference core/attention.py:60-70).  With the 16-bit weight shadows of the three layers back to back (they
    are: `WeightShadows` lays them out that way) the projections are ONE library GEMM [rows, d] x [d, 3 d] forward; q and
    k are rotated out of its output by one launch, v stays a column slice of it.  Backward: the attention backward
    writes dq | dk | dv into one [rows, 3 d] buffer (`joint_grads`), dq and dk are rotated back in place by one launch
    that also accumulates their column sums (the q / k bias gradients), and the input and weight gradients are one
    GEMM each (89 / 78 / 91 us against 114 / 123 / 124 us for three of each at [65536, 512]).  Without adjacent
    shadows or a joint gradient buffer it is the three-GEMM form of `_Linear3Fn` around the same rotation launch."""

    @staticmethod
    def forward(ctx, x: Tensor, dtype: torch.dtype, shadows, cos: Tensor, sin: Tensor, *wb):
        x16 = x.to(dtype)
        w16s, b16s = [], []
        for i in range(3):
            w, b, w16, b16 = wb[4 * i:4 * i + 4]
            if w16 is None:
                w16, b16 = w.to(dtype), b.to(dtype)
            w16s.append(w16)
            b16s.append(b16)
        d_out, d_in = w16s[0].shape
        joint = _adjacent(w16s, d_out * d_in) and _adjacent(b16s, d_out) and d_out % 8 == 0
        if joint:
            w_cat = torch.as_strided(w16s[0], (3 * d_out, d_in), (d_in, 1))
            b_cat = torch.as_strided(b16s[0], (3 * d_out,), (1,))
            qkv = F.linear(x16, w_cat, b_cat)
            q0, k0, v = qkv[..., :d_out], qkv[..., d_out:2 * d_out], qkv[..., 2 * d_out:]
        else:
            q0, k0, v = (F.linear(x16, w16, b16) for w16, b16 in zip(w16s, b16s))
        q, k, _, _ = rotary_pair(q0, k0, cos, sin)
        ctx.save_for_backward(x16, cos, sin, *w16s)
        ctx.in_dtypes = (x.dtype, wb[0].dtype, wb[1].dtype)
        ctx.shadows, ctx.epoch = shadows, (shadows.epoch if shadows is not None else 0)
        return q, k, v

    @staticmethod
    def backward(ctx, gq, gk, gv):
        x16, cos, sin, *w16s = ctx.saved_tensors
        if ctx.shadows is not None and ctx.shadows.epoch != ctx.epoch:
            raise RuntimeError("a weight needed for this backward pass was modified (optimizer step?) after the forward "
                               "pass that used its 16-bit shadow copy")
        xd, wd, bd = ctx.in_dtypes
        need = ctx.needs_input_grad
        x2 = x16.reshape(-1, x16.shape[-1])
        d_out, d_in = w16s[0].shape
        # ---- one [rows, 3 d] gradient buffer (attention backward with joint_grads) and adjacent weights: one GEMM each
        # (adjacency is checked on the tensors at hand: under activation checkpointing the saved weights are those of the
        #  recomputed forward, which runs outside `WeightShadows.step()` and casts each weight separately)
        if (_adjacent(w16s, d_out * d_in) and gq is not None and gk is not None and gv is not None and all(need[5 + i] for i in range(0, 12, 4))
                and all(need[6 + i] for i in range(0, 12, 4)) and not (gq.requires_grad or gk.requires_grad or gv.requires_grad)):
            ld = _row_stride(gq)
            if (ld == 3 * d_out and _row_stride(gk) == ld and _row_stride(gv) == ld and gq.dtype == gk.dtype == gv.dtype == x16.dtype
                    and gk.data_ptr() == gq.data_ptr() + d_out * gq.element_size()
                    and gv.data_ptr() == gk.data_ptr() + d_out * gq.element_size()):
                rows = x2.shape[0]
                _, _, sum_q, sum_k = rotary_pair(gq, gk, cos, sin, conj=True, want_colsum=True, inplace=True)
                dqkv = torch.as_strided(gq, (rows, 3 * d_out), (ld, 1))
                w_cat = torch.as_strided(w16s[0], (3 * d_out, d_in), (d_in, 1))
                dx = torch.mm(dqkv, w_cat).view(x16.shape).to(xd) if need[0] else None
                dw = torch.mm(dqkv.t(), x2, out_dtype=torch.float32)
                sum_v = colsum(torch.as_strided(gv, (rows, d_out), (ld, 1)))
                grads = []
                for i, sm in enumerate((sum_q, sum_k, sum_v)):
                    grads += [dw[i * d_out:(i + 1) * d_out].to(wd), sm.to(bd), None, None]
                return (dx, None, None, None, None, *grads)
        # ---- three GEMM pairs
        sums = [None, None, None]
        gs = [gq, gk, gv]
        if gq is not None and gk is not None:
            want = bool(need[6] or need[10])
            gs[0], gs[1], sums[0], sums[1] = rotary_pair(gq, gk, cos, sin, conj=True, want_colsum=want)
        else:                                      # (one of q / k unused downstream: rotate what is there with a zero partner)
            for i in (0, 1):
                if gs[i] is not None:
                    gs[i] = rotary_pair(gs[i], torch.zeros_like(gs[i]), cos, sin, conj=True)[0]
        dx = None
        grads = []
        for i, (g, w16) in enumerate(zip(gs, w16s)):
            dw = db = None
            if g is not None:
                g2 = g.reshape(-1, w16.shape[0])
                if not g2.is_contiguous():
                    g2 = g2.contiguous()
                if need[0]:
                    if dx is None:
                        dx = torch.mm(g2, w16)
                    else:
                        dx.addmm_(g2, w16)
                if need[5 + 4 * i]:
                    dw = torch.mm(g2.t(), x2, out_dtype=torch.float32).to(wd)
                if need[6 + 4 * i]:
                    db = (sums[i] if sums[i] is not None else colsum(g2)).to(bd)
            grads += [dw, db, None, None]
        return (dx.view(x16.shape).to(xd) if dx is not None else None, None, None, None, None, *grads)


def qkv_rotary(x: Tensor, lin_q: 'Linear', lin_k: 'Linear', lin_v: 'Linear', cos: Tensor, sin: Tensor):
    """(rotary(lin_q(x)), rotary(lin_k(x)), lin_v(x)) as one autograd node, or None where the fused form does not apply
    (the caller then runs the separate ops)."""
    if not all(m._fused_ok(x) for m in (lin_q, lin_k, lin_v)):
        return None
    dtype = torch.get_autocast_dtype('cuda')
    if dtype not in (torch.bfloat16, torch.float16) or cos.dtype not in (dtype, torch.float32) or x.shape[-1] % 8:
        return None
    args = []
    shs = [m._active_shadow(dtype) for m in (lin_q, lin_k, lin_v)]
    use = all(sh is not None for sh in shs)
    for m, sh in zip((lin_q, lin_k, lin_v), shs):
        args += [m.weight, m.bias, sh[1] if use else None, sh[2] if use else None]
    return _QkvRotaryFn.apply(x, dtype, shs[0][0] if use else None, cos, sin, *args)


def linear3(x: Tensor, lin0: 'Linear', lin1: 'Linear', lin2: 'Linear'):
    """(lin0(x), lin1(x), lin2(x)) -- the q / k / v projections of self-attention."""
    if all(m._fused_ok(x) for m in (lin0, lin1, lin2)):
        dtype = torch.get_autocast_dtype('cuda')
        if dtype in (torch.bfloat16, torch.float16):
            args, shadows = [], None
            shs = [m._active_shadow(dtype) for m in (lin0, lin1, lin2)]
            use = all(sh is not None for sh in shs)
            for m, sh in zip((lin0, lin1, lin2), shs):
                args += [m.weight, m.bias, sh[1] if use else None, sh[2] if use else None]
            if use:
                shadows = shs[0][0]
            return _Linear3Fn.apply(x, dtype, shadows, *args)
    return lin0(x), lin1(x), lin2(x)


class Linear(nn.Linear):
    _shadow = None          # (WeightShadows, weight view, bias view | None), set by WeightShadows

    def _fused_ok(self, x: Tensor) -> bool:
        return bool(N.FUSED_EXTRAS and x.is_cuda and torch.is_autocast_enabled('cuda') and torch.is_grad_enabled()
                    and self.out_features % 8 == 0 and x.numel() >= _MIN_ROWS * self.in_features
                    and (x.requires_grad or self.weight.requires_grad))

    def _active_shadow(self, dtype):
        sh = self._shadow
        return sh if sh is not None and sh[0] is WeightShadows.ACTIVE and sh[1].dtype == dtype else None

    def forward(self, x: Tensor) -> Tensor:
        if self._fused_ok(x):
            dtype = torch.get_autocast_dtype('cuda')
            if dtype in (torch.bfloat16, torch.float16):
                sh = self._active_shadow(dtype)
                if sh is not None:
                    return _LinearFn.apply(x, self.weight, self.bias, dtype, sh[1], sh[2], sh[0])
                return _LinearFn.apply(x, self.weight, self.bias, dtype, None, None, None)
        return F.linear(x, self.weight, self.bias)


class WeightShadows:
    """16-bit copies of the parameters of every `Linear` below `root`, refreshed together once per training step.

        with shadows.step():            # refresh (a few launches), then let the Linear layers use the copies
            loss = forward(...)
        loss.backward()                 # the saved views stay valid until the next refresh changes their values

    The copies are only read while `step()` is open and they are rebuilt from the fp32 parameters every time it is
    entered, so any update of the parameters between steps (optimizers, `load_state_dict`, `.data` writes) is picked
    up.  A backward pass that runs after the parameters changed AND the copies were refreshed raises, like autograd
    does for a modified fp32 weight."""

    ACTIVE: Optional['WeightShadows'] = None
    ENABLED = True              # tests switch the mechanism off to compare against the per-weight casts

    def __init__(self, root: nn.Module):
        self.root = root
        self.epoch = 0
        self._params: List[Tensor] = []
        self._views: List[Tensor] = []
        self._flat = None
        self._key = None
        self._versions = None
        self._dst, self._src = _PtrList(), _PtrList()

    def _collect(self, dtype: torch.dtype):
        mods = [m for m in self.root.modules() if isinstance(m, Linear) and m.weight.is_cuda
                and m.weight.dtype == torch.float32 and m.weight.is_contiguous()]
        params, index = [], {}
        # all weights first (module order), then all biases: the q / k / v weights of an attention layer -- consecutive
        # modules of equal shape -- sit back to back, and so do their biases, which lets `_QkvRotaryFn` run the three
        # projections as one GEMM on a [3 d, d] view
        for pick in (lambda m: m.weight, lambda m: m.bias):
            for m in mods:
                t = pick(m)
                if t is not None and id(t) not in index and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous():
                    index[id(t)] = len(params)
                    params.append(t)
        offsets, total = [], 0
        for t in params:
            offsets.append(total)
            total += (t.numel() + 7) // 8 * 8                   # 16-byte aligned views
        dev = params[0].device if params else torch.device('cuda')
        self._flat = torch.empty(max(total, 8), dtype=dtype, device=dev)
        self._params = params
        self._views = [self._flat[o:o + t.numel()].view(t.shape) for o, t in zip(offsets, params)]
        for m in self.root.modules():
            if isinstance(m, Linear):
                m._shadow = None
        for m in mods:
            if id(m.weight) not in index or (m.bias is not None and id(m.bias) not in index):
                continue
            m._shadow = (self, self._views[index[id(m.weight)]],
                         self._views[index[id(m.bias)]] if m.bias is not None else None)

    def refresh(self):
        dtype = torch.get_autocast_dtype('cuda')
        key = (dtype,) + tuple((id(m), m.weight.data_ptr(), None if m.bias is None else m.bias.data_ptr())
                               for m in self.root.modules() if isinstance(m, Linear))
        if key != self._key:
            self._collect(dtype)
            self._key = key
        if not self._params:
            return
        d, s = self._dst.update(self._views), self._src.update(self._params)
        N.check(N.lib.svae_multi_tensor_cast(d.n, d.ptrs, s.ptrs, s.numel, N.svae_dtype(dtype),
                                             N.current_stream(self._params[0].device)), 'svae_multi_tensor_cast')
        versions = sum(t._version for t in self._params)
        if versions != self._versions:
            self._versions = versions
            self.epoch += 1

    @contextmanager
    def step(self):
        self.refresh()
        previous, WeightShadows.ACTIVE = WeightShadows.ACTIVE, self
        try:
            yield self
        finally:
            WeightShadows.ACTIVE = previous
