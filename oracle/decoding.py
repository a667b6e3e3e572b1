"""Oracle (test infrastructure): token-by-token decoding of `TransformerVAE.sample`.

Reference: sparse_vae/core/generation.py:40-72 (`GenerationState.process_logits`: repetition penalty, temperature,
top-k / nucleus filtering, multinomial draw) and the KV cache of sparse_vae/core/attention.py:107-142 (keys kept for a
sparse layer: the global first block plus a sliding window of `window_size` blocks).

  nucleus_keep_reference   the reference's own rule, op for op: sort descending, softmax, drop positions whose
                           INCLUSIVE cumulative mass exceeds top_p, always keep the first (generation.py:55-62)
  nucleus_weights          the same set stated without a sort (what csrc/sampling.cu computes): most likely values
                           first while their mass stays <= top_p; tokens tying at the boundary admitted in index order
                           while they fit; never empty.  Returns exp(x - max) for kept tokens, 0 elsewhere (float64).
  inverse_cdf_token        the token an inverse-CDF draw over those weights (index order) returns for a uniform u
  visible_positions        the key positions a query at position p may attend: rows of the causal include_cls layout
                           (oracle/layout.py) restricted to keys <= p -- what both the reference's sliding cache and the
                           ring cache of csrc/decode_attn.cu must expose
"""
from __future__ import annotations

import torch

from . import layout as olayout


def nucleus_keep_reference(logits_row: torch.Tensor, top_p: float) -> torch.Tensor:
    x = logits_row.float()
    sorted_x, order = x.sort(descending=True)
    probs = sorted_x.softmax(-1)
    tail = probs.cumsum(-1) > top_p
    tail[0] = False
    keep = torch.zeros(x.numel(), dtype=torch.bool)
    keep[order[~tail]] = True
    return keep


def nucleus_weights(x64: torch.Tensor, top_p: float) -> torch.Tensor:
    e = torch.exp(x64 - x64.max())
    budget = top_p * e.sum()
    keep = torch.zeros_like(e, dtype=torch.bool)
    tail = 0.0
    for value in torch.unique(x64).flip(0).tolist():
        group = x64 == value
        mass = e[group].sum().item()
        if tail + mass <= budget:
            keep |= group
            tail += mass
            continue
        each = e[group][0].item()
        n = int((budget.item() - tail) // each)
        if tail == 0.0:
            n = max(n, 1)
        keep[group.nonzero().flatten()[:n]] = True
        break
    return torch.where(keep, e, torch.zeros_like(e))


def inverse_cdf_token(weights64: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    cdf = weights64.cumsum(0)
    last_kept = int((weights64 > 0).nonzero()[-1])            # u * total rounding up to the total: the last kept token
    return torch.searchsorted(cdf, u.double() * cdf[-1], right=True).clamp_(max=last_kept)


def visible_positions(p: int, window: int, block: int = 32):
    nb = p // block + 1
    lay = olayout.layout_2d(nb, window, causal=True, include_cls=True)
    return {q for q in range(p + 1) if lay[p // block, q // block]}
