"""Synthetic token batches with the schema of the reference's collate function
(sparse_vae/text_data_module.py:194-210): {'token_ids': PaddedTensor, 'num_tokens', 'num_bytes'}; ids uniform in
[3, 32768), position 0 = [CLS] (1), last real token = [SEP] (2), padding id 0."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .core.padded_tensor import PaddedTensor

CLS, SEP, PAD = 1, 2, 0


def synthetic_tokens(batch: int, seq_len: int, seed: int = 7295, lengths: Optional[Sequence[int]] = None,
                     dtype=torch.int16, pin: bool = False) -> dict:
    """Host-side batch (int16 ids like the reference's collate); `lengths` None = every sequence is full."""
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(3, 2 ** 15, (batch, seq_len), generator=g, dtype=torch.int64)
    tok[:, 0] = CLS
    n = torch.full((batch,), seq_len, dtype=torch.int64) if lengths is None else torch.as_tensor(lengths, dtype=torch.int64)
    for b in range(batch):
        tok[b, n[b] - 1] = SEP
        tok[b, n[b]:] = PAD
    tok = tok.to(dtype)
    if pin:
        tok, n = tok.pin_memory(), n.pin_memory()
    return {'token_ids': tok, 'num_tokens': n, 'num_bytes': 4 * n}


def to_device(host_batch: dict, device, non_blocking: bool = True) -> dict:
    tok = host_batch['token_ids'].to(device, non_blocking=non_blocking)
    n = host_batch['num_tokens'].to(device, non_blocking=non_blocking)
    return {'token_ids': PaddedTensor.from_raw(tok), 'num_tokens': n, 'num_bytes': 4 * n}
