"""Oracle (test infrastructure): latent bottleneck = ConditionalGaussian + rsample + KL reduction.

Reference path:
  core/conditional_gaussian.py:18-30   mu, logvar = Linear(x).chunk(2,-1); var = exp(logvar);
                                       Normal(mu, sqrt(var)); kl = 0.5*(mu^2 + var - logvar - 1)
  core/continuous_autoencoder.py:42-52 z = q.rsample(); raw_kl = kl.flatten(1).sum(-1);
                                       kl = (raw_kl / token_counts).mean()
  torch.distributions.Normal.rsample   eps = empty(shape, dtype=loc.dtype).normal_(); z = loc + eps*scale

The CUDA `normal_` stream is restated from torch's ATen headers (third-party, pinned by the image:
torch 2.11.0): include/ATen/native/cuda/DistributionTemplates.h:50-92 (grid/offset policy and the
grid-stride loop) and curand's Philox4x32-10 + Box-Muller (curand_philox4x32_x.h, curand_normal.h:70-87).

The CPU restatement of eps is NOT bit-exact with the GPU: curand evaluates Box-Muller with the fast
`__sincosf` intrinsic and device `logf`.  It agrees to ~1e-6 absolute in fp32, i.e. at most one unit
in the last place after rounding to bf16/fp16 on a small fraction of elements.  Bit-exactness of z is
instead asserted on the GPU against torch's own `Normal.rsample` with the same generator state
(tests/test_gpu_bottleneck.py).
"""
from __future__ import annotations

import numpy as np
import torch

PHILOX_W32_0 = np.uint32(0x9E3779B9)
PHILOX_W32_1 = np.uint32(0xBB67AE85)
PHILOX_M4x32_0 = np.uint64(0xD2511F53)
PHILOX_M4x32_1 = np.uint64(0xCD9E8D57)
CURAND_2POW32_INV = np.float32(2.3283064e-10)
CURAND_2POW32_INV_2PI = np.float32(np.float32(2.3283064e-10) * np.float32(6.2831855))

BLOCK = 256          # block_size_bound, DistributionTemplates.h
UNROLL = 4           # sizeof(float4)/sizeof(float)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: [...,4] uint32, key: [...,2] uint32 -> [...,4] uint32 (curand_Philox4x32_10)."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint32).copy()
    k1 = key[..., 1].astype(np.uint32).copy()
    mask = np.uint64(0xFFFFFFFF)
    for rnd in range(10):
        p0 = PHILOX_M4x32_0 * c[0]
        p1 = PHILOX_M4x32_1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        n0 = hi1 ^ c[1] ^ k0.astype(np.uint64)
        n1 = lo1
        n2 = hi0 ^ c[3] ^ k1.astype(np.uint64)
        n3 = lo0
        c = [n0 & mask, n1, n2 & mask, n3]
        if rnd < 9:
            k0 = (k0 + PHILOX_W32_0).astype(np.uint32)
            k1 = (k1 + PHILOX_W32_1).astype(np.uint32)
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def box_muller(x: np.ndarray, y: np.ndarray):
    """curand_normal.h:70-87 in fp32."""
    u = x.astype(np.float32) * CURAND_2POW32_INV + np.float32(CURAND_2POW32_INV / np.float32(2))
    v = y.astype(np.float32) * CURAND_2POW32_INV_2PI + np.float32(CURAND_2POW32_INV_2PI / np.float32(2))
    s = np.sqrt(np.float32(-2.0) * np.log(u, dtype=np.float32), dtype=np.float32)
    return (np.sin(v, dtype=np.float32) * s).astype(np.float32), (np.cos(v, dtype=np.float32) * s).astype(np.float32)


def execution_policy(numel: int, sm_count: int, max_threads_per_sm: int = 2048):
    """(counter_offset, grid) of calc_execution_policy, DistributionTemplates.h:50-63."""
    grid = (numel + BLOCK - 1) // BLOCK
    grid = min(sm_count * (max_threads_per_sm // BLOCK), grid)
    counter_offset = ((numel - 1) // (BLOCK * grid * UNROLL) + 1) * 4
    return counter_offset, grid


def standard_normal_like_cuda(numel: int, seed: int, offset: int, sm_count: int = 148,
                              max_threads_per_sm: int = 2048) -> np.ndarray:
    """fp32 standard normals in the element order torch's CUDA `normal_` would produce.

    Element li is produced by thread `li mod T` (T = 256*grid) on its `(li div T) div 4`-th
    curand_normal4 call, component `(li div T) mod 4` (DistributionTemplates.h:66-92).
    curand_init(seed, subsequence=thread, offset): counter = (offset/4 + call, 0, thread_lo, thread_hi).
    """
    assert offset % 4 == 0, "torch always advances the Philox offset in multiples of 4"
    _, grid = execution_policy(numel, sm_count, max_threads_per_sm)
    T = BLOCK * grid
    li = np.arange(numel, dtype=np.int64)
    thread = li % T
    k = li // T
    call, comp = k // 4, k % 4
    ctr64 = np.uint64(offset // 4) + call.astype(np.uint64)
    ctr = np.stack([
        (ctr64 & np.uint64(0xFFFFFFFF)).astype(np.uint32),
        (ctr64 >> np.uint64(32)).astype(np.uint32),
        (thread.astype(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32),
        (thread.astype(np.uint64) >> np.uint64(32)).astype(np.uint32),
    ], axis=-1)
    key = np.empty((numel, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(ctr, key)
    n0, n1 = box_muller(r[:, 0], r[:, 1])
    n2, n3 = box_muller(r[:, 2], r[:, 3])
    allc = np.stack([n0, n1, n2, n3], axis=-1)
    return allc[np.arange(numel), comp]


def bottleneck_forward(mulogvar: torch.Tensor, token_counts: torch.Tensor, eps: torch.Tensor):
    """mulogvar [B, ..., 2D] (the Linear output), eps like mu (already rounded to mu.dtype).

    Returns dict(z, sigma, kl_elem, raw_kl, kl) in fp32, following conditional_gaussian.py:19-27 and
    continuous_autoencoder.py:43-47 with CUDA-autocast promotion (exp/pow/sqrt in fp32).
    """
    mu, logvar = mulogvar.chunk(2, dim=-1)
    mu32, logvar32 = mu.float(), logvar.float()
    var = logvar32.exp()
    sigma = var.sqrt()
    z = mu32 + eps.float() * sigma
    kl_elem = 0.5 * (mu32 ** 2 + var - logvar32 - 1.0)
    raw_kl = kl_elem.flatten(1).sum(dim=-1)
    kl = (raw_kl / token_counts.float()).mean()
    return dict(z=z, sigma=sigma, kl_elem=kl_elem, raw_kl=raw_kl, kl=kl)


def bottleneck_backward(mulogvar: torch.Tensor, token_counts: torch.Tensor, eps: torch.Tensor,
                        dz: torch.Tensor, dkl: float):
    """Analytic gradient w.r.t. mulogvar:
    d/dmu = dz + g*mu ; d/dlogvar = dz*0.5*eps*sigma + g*0.5*(var - 1), g = dkl/(B*token_counts[b])."""
    mu, logvar = mulogvar.chunk(2, dim=-1)
    mu32, logvar32 = mu.float(), logvar.float()
    var = logvar32.exp()
    sigma = var.sqrt()
    B = mulogvar.shape[0]
    g = (dkl / (B * token_counts.float())).view(B, *([1] * (mu.ndim - 1)))
    dmu = dz.float() + g * mu32
    dlogvar = dz.float() * 0.5 * eps.float() * sigma + g * 0.5 * (var - 1.0)
    return torch.cat([dmu, dlogvar], dim=-1)
