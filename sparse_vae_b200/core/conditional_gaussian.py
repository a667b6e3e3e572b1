"""`ConditionalGaussian` with the fused bottleneck kernel (reference: sparse_vae/core/conditional_gaussian.py:6-30).

Same constructor, parameters (`linear`) and return types as the reference.  The element-wise tail of the reference
(`chunk -> exp -> sqrt -> pow/add/sub/mul`, then `Normal.rsample` and the KL reductions in
`ContinuousVAE.sample_z`, core/continuous_autoencoder.py:42-52) is one launch of `svae_bottleneck_fwd`; gradients
come from `svae_bottleneck_bwd`, which regenerates eps from the saved Philox (seed, offset).
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch
from torch import nn, Tensor
from torch.distributions import Normal

from .. import _native as N

_workspaces = {}


def _workspace(device: torch.device) -> Tensor:
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = _workspaces[key] = torch.zeros(N.BOTTLENECK_WORKSPACE_BYTES, dtype=torch.uint8, device=device)
    return ws


def _device_geometry(device: torch.device):
    props = torch.cuda.get_device_properties(device)
    return props.multi_processor_count, props.max_threads_per_multi_processor


def philox_reserve(device: torch.device, rows: int, latent: int, generator: Optional[torch.Generator] = None):
    """Takes (seed, offset) from the CUDA generator and advances it exactly as `normal_()` on a
    [rows*latent] tensor would (ATen calc_execution_policy), so later draws stay in step with the reference."""
    sms, tpm = _device_geometry(device)
    increment = int(N.lib.svae_bottleneck_philox_increment(rows, latent, sms, tpm))
    from .graph_step import StepPhilox
    if StepPhilox.active is not None and generator is None:
        # a training step is being captured: {seed, base offset} are read from device memory at replay time and this
        # draw sits `delta` past the base (third element: the device pointer)
        philox_dev, delta = StepPhilox.active.reserve(increment)
        return 0, delta, philox_dev
    gen = generator if generator is not None else torch.cuda.default_generators[
        device.index if device.index is not None else torch.cuda.current_device()]
    seed, offset = gen.initial_seed(), gen.get_offset()
    gen.set_offset(offset + increment)
    return seed, offset


class _BottleneckFn(torch.autograd.Function):
    """(mulogvar[rows, 2D], counts[rows]) -> z, sigma, kl_elem [rows, D] fp32, raw_kl[rows], kl[] (mean of raw_kl/counts)."""

    @staticmethod
    def forward(ctx, mulogvar: Tensor, counts: Tensor, seed: int, offset: int, philox_dev: int = 0):
        if not mulogvar.is_cuda:
            raise ValueError("Only GPU devices are supported for now")
        rows, two_d = mulogvar.shape
        D = two_d // 2
        if mulogvar.stride(1) != 1:
            mulogvar = mulogvar.contiguous()
        dev = mulogvar.device
        sms, tpm = _device_geometry(dev)
        z = torch.empty(rows, D, dtype=torch.float32, device=dev)
        sigma, kl_elem = torch.empty_like(z), torch.empty_like(z)
        raw_kl = torch.empty(rows, dtype=torch.float32, device=dev)
        kl = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            N.check(N.lib.svae_bottleneck_fwd_g(
                N.ptr(mulogvar), mulogvar.stride(0), N.svae_dtype(mulogvar.dtype), N.ptr(counts), rows, D, seed, offset,
                philox_dev, sms, tpm, N.ptr(z), N.ptr(sigma), N.ptr(kl_elem), N.ptr(raw_kl), N.ptr(kl), N.ptr(_workspace(dev)),
                N.current_stream(dev)), 'svae_bottleneck_fwd')
        ctx.save_for_backward(mulogvar, counts)
        ctx.philox = (seed, offset, philox_dev)
        return z, sigma, kl_elem, raw_kl, kl

    @staticmethod
    def backward(ctx, dz, dsigma, dkl_elem, draw_kl, dkl):
        mulogvar, counts = ctx.saved_tensors
        rows, two_d = mulogvar.shape
        D = two_d // 2
        dev = mulogvar.device
        sms, tpm = _device_geometry(dev)
        seed, offset, philox_dev = ctx.philox

        def f32(t):
            return None if t is None else t.to(torch.float32).contiguous()

        dz, dsigma, dkl_elem, draw_kl, dkl = map(f32, (dz, dsigma, dkl_elem, draw_kl, dkl))
        grad = torch.empty(rows, two_d, dtype=mulogvar.dtype, device=dev)
        with torch.cuda.device(dev):
            N.check(N.lib.svae_bottleneck_bwd_g(
                N.ptr(mulogvar), mulogvar.stride(0), N.svae_dtype(mulogvar.dtype), N.ptr(counts), rows, D, seed, offset,
                philox_dev, sms, tpm, N.ptr(dz), N.ptr(dsigma), N.ptr(dkl_elem), N.ptr(draw_kl), N.ptr(dkl), N.ptr(grad),
                grad.stride(0), N.current_stream(dev)), 'svae_bottleneck_bwd')
        return grad, None, None, None, None


def fused_bottleneck(mulogvar: Tensor, token_counts: Optional[Tensor], philox: Optional[Tuple[int, int]] = None,
                     generator: Optional[torch.Generator] = None):
    """mulogvar [B, ..., 2D] (the Linear output) -> dict(z, sigma, kl_elem [B, ..., D] fp32, raw_kl [B], kl []).

    raw_kl sums over everything but the batch dimension and kl = mean(raw_kl / token_counts), as in
    ContinuousVAE.sample_z (reference core/continuous_autoencoder.py:46-47).
    """
    mulogvar = mulogvar.as_subclass(Tensor) if type(mulogvar) is not Tensor else mulogvar
    if not mulogvar.is_cuda:
        raise ValueError("Only GPU devices are supported for now")
    lead = mulogvar.shape[:-1]
    D = mulogvar.shape[-1] // 2
    B = lead[0]
    flat = mulogvar.reshape(-1, 2 * D)
    rows = flat.shape[0]
    per_batch = rows // B
    if token_counts is None:
        counts = torch.ones(rows, dtype=torch.int64, device=mulogvar.device)
    else:
        counts = token_counts.as_subclass(Tensor).to(device=mulogvar.device, dtype=torch.int64).reshape(B)
        if per_batch > 1:
            counts = counts.repeat_interleave(per_batch)
        counts = counts.contiguous()
    ph = philox if philox is not None else philox_reserve(mulogvar.device, rows, D, generator)
    z, sigma, kl_elem, raw_kl, kl = _BottleneckFn.apply(flat, counts, *ph)
    if per_batch > 1:
        raw_kl = raw_kl.reshape(B, per_batch).sum(dim=-1)
        kl = kl * per_batch
    return dict(z=z.reshape(*lead, D), sigma=sigma.reshape(*lead, D), kl_elem=kl_elem.reshape(*lead, D),
                raw_kl=raw_kl, kl=kl)


class ConditionalGaussian(nn.Module):
    def __init__(self, in_features: int, out_features: int, zero_initialized: bool = False, bias: bool = True):
        super(ConditionalGaussian, self).__init__()

        linear = nn.Linear(in_features, out_features * 2, bias=bias)
        if zero_initialized:
            linear.weight.data.zero_()
            if bias:
                linear.bias.data.zero_()

        self.linear = linear

    def forward(self, x: Tensor, get_kl: bool = False) -> Union[Normal, Tuple[Normal, Tensor]]:
        mulogvar = self.linear(x)
        mu = mulogvar.chunk(2, dim=-1)[0]
        # sigma = sqrt(exp(logvar)) and the element-wise KL from the fused kernel; its z output is unused here, so
        # a fixed Philox position is passed and the generator is not advanced.
        fused = fused_bottleneck(mulogvar, None, philox=(0, 0))
        # No parameter validation, like the reference (sigma == 0 must give an infinite KL, not an exception)
        gaussian = Normal(loc=mu, scale=fused['sigma'], validate_args=False)
        return (gaussian, fused['kl_elem']) if get_kl else gaussian

    def sample(self, x: Tensor, token_counts: Tensor, generator: Optional[torch.Generator] = None):
        """Fused `forward(get_kl=True)` + `rsample()` + KL reductions: returns (z, kl, raw_kl, Normal)."""
        mulogvar = self.linear(x)
        mu = mulogvar.chunk(2, dim=-1)[0]
        fused = fused_bottleneck(mulogvar, token_counts, generator=generator)
        gaussian = Normal(loc=mu, scale=fused['sigma'], validate_args=False)
        return fused['z'], fused['kl'], fused['raw_kl'], gaussian
