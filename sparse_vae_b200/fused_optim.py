"""Host side of the fused multi-tensor optimizer kernels (csrc/optim.cu): gradient-norm clipping and the RAdam
update over every parameter tensor of the model in a handful of launches.

Reference call sites: `LanguageModel.on_after_backward` (sparse_vae/core/language_model.py:120-122) and
`RAdam.step` (sparse_vae/core/rectified_adam.py:15-88).  CUDA fp32 tensors only; callers keep the plain torch
implementation for CPU tensors (the gloo data-parallel tests run there).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch
from torch import Tensor

from . import _native as N


class _PtrList:
    """ctypes arrays of device pointers / element counts for a list of tensors, rebuilt only when a pointer moves."""

    def __init__(self):
        self.key = None
        self.ptrs = None
        self.numel = None
        self.n = 0

    def update(self, tensors: Sequence[Tensor]):
        key = tuple(t.data_ptr() for t in tensors)
        if key != self.key:
            self.key = key
            self.n = len(tensors)
            self.ptrs = (C.c_void_p * self.n)(*key)
            self.numel = (C.c_int64 * self.n)(*(t.numel() for t in tensors))
        return self


def _check_fp32_cuda(tensors: Sequence[Tensor], what: str):
    for t in tensors:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError(f"{what}: fused optimizer kernels need contiguous CUDA float32 tensors, got "
                             f"{t.dtype} on {t.device} (contiguous={t.is_contiguous()})")


class FusedGradClipper:
    """`torch.nn.utils.clip_grad_norm_` for a fixed parameter set: returns the total norm as a 0-d device tensor
    (no host synchronisation) and scales every gradient in place by min(1, max_norm / (norm + 1e-6))."""

    def __init__(self):
        self._list = _PtrList()
        self._partials = None
        self._out = None

    def __call__(self, grads: List[Tensor], max_norm: float) -> Tensor:
        if not grads:
            return torch.zeros(())
        _check_fp32_cuda(grads, 'clip_grad_norm')
        dev = grads[0].device
        pl = self._list.update(grads)
        chunks = N.lib.svae_multi_tensor_chunks(pl.n, pl.numel)
        if self._partials is None or self._partials.numel() < chunks or self._partials.device != dev:
            self._partials = torch.empty(max(chunks, 1), device=dev, dtype=torch.float32)
        out = torch.empty(2, device=dev, dtype=torch.float32)
        N.check(N.lib.svae_clip_grad_norm(pl.n, pl.ptrs, pl.numel, float(max_norm), self._partials.data_ptr(),
                                          self._partials.numel(), out.data_ptr(), N.current_stream(dev)),
                'svae_clip_grad_norm')
        return out[0]


class FusedRAdamStep:
    def __init__(self):
        self._p, self._g, self._m, self._v = _PtrList(), _PtrList(), _PtrList(), _PtrList()

    def __call__(self, params: List[Tensor], grads: List[Tensor], exp_avg: List[Tensor], exp_avg_sq: List[Tensor],
                 lr: float, beta1: float, beta2: float, eps: float, weight_decay: float, step: int, args_dev: int = 0):
        _check_fp32_cuda(params, 'RAdam params')
        _check_fp32_cuda(grads, 'RAdam grads')
        dev = params[0].device
        p, g = self._p.update(params), self._g.update(grads)
        m, v = self._m.update(exp_avg), self._v.update(exp_avg_sq)
        if args_dev:        # a captured step: the scalars are read from device memory at replay time (core/graph_step.py)
            N.check(N.lib.svae_radam_step_g(p.n, p.ptrs, g.ptrs, m.ptrs, v.ptrs, p.numel, args_dev, N.current_stream(dev)),
                    'svae_radam_step_g')
        else:
            N.check(N.lib.svae_radam_step(p.n, p.ptrs, g.ptrs, m.ptrs, v.ptrs, p.numel, float(lr), float(beta1), float(beta2),
                                          float(eps), float(weight_decay), int(step), N.current_stream(dev)),
                    'svae_radam_step')
        # the kernel wrote through raw pointers: tell autograd (and anything keyed on tensor versions) about it
        torch.autograd.graph.increment_version(params)


class FusedScaleCopy:
    """dst[i] = src[i] * scale over two equally shaped lists of CUDA fp32 tensors, a handful of launches in total."""

    def __init__(self):
        self._d, self._s = _PtrList(), _PtrList()

    def __call__(self, dst: List[Tensor], src: List[Tensor], scale: float = 1.0):
        if not dst:
            return
        _check_fp32_cuda(dst, 'scale_copy dst')
        _check_fp32_cuda(src, 'scale_copy src')
        d, s = self._d.update(dst), self._s.update(src)
        N.check(N.lib.svae_multi_tensor_scale_copy(d.n, d.ptrs, s.ptrs, d.numel, float(scale),
                                                   N.current_stream(dst[0].device)), 'svae_multi_tensor_scale_copy')
