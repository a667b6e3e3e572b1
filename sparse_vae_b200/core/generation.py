"""Decode-time sampling state (reference surface: sparse_vae/core/generation.py): top-k / nucleus /
temperature sampling with a repetition penalty over the last 512 generated tokens."""
from __future__ import annotations

from dataclasses import dataclass, InitVar

import torch
from torch import Tensor


@dataclass
class GenerationState:
    max_length: InitVar[int]
    batch_size: InitVar[int]
    start_token: int
    end_token: int
    device: InitVar[torch.device]
    dtype: InitVar[torch.dtype] = torch.long

    top_k: int = 0
    top_p: float = 0.9
    temperature: float = 1.0
    repetition_penalty: float = 1.2

    def __post_init__(self, max_length: int, batch_size: int, device: torch.device, dtype: torch.dtype):
        self.output_ids = torch.zeros(batch_size, max_length, device=device, dtype=dtype)
        self.output_ids[:, 0] = self.start_token
        self.live_sample_mask = torch.ones(batch_size, device=device, dtype=torch.bool)
        self.current_index = 1

    def prev_tokens(self) -> Tensor:
        return self.output_ids[self.live_sample_mask, self.current_index - 1, None]

    def process_logits(self, logits: Tensor):
        ids = None
        if self.repetition_penalty > 1.0:
            start = max(self.current_index - 512, 0)
            seen = self.output_ids[self.live_sample_mask, start:self.current_index]
            seen_logits = logits.gather(dim=-1, index=seen)
            seen_logits = torch.where(seen_logits < 0.0, seen_logits * self.repetition_penalty,
                                      seen_logits / self.repetition_penalty)
            logits.scatter_(dim=-1, index=seen, src=seen_logits)

        if self.temperature <= 0.0 or self.top_k == 1:          # greedy
            logits, ids = logits.max(dim=-1, keepdim=True)
        else:
            logits = logits / self.temperature
            if self.top_k > 0:
                logits, ids = logits.topk(k=max(self.top_k, 1), sorted=False)
            if self.top_p < 1.0:                                # nucleus: drop the tail beyond cumulative mass top_p
                logits, order = logits.sort(descending=True)
                ids = order if ids is None else ids.gather(dim=-1, index=order)
                probs = logits.softmax(dim=-1)
                tail = probs.cumsum(dim=-1) > self.top_p
                tail[..., :1] = False
                probs[tail] = 0.0
            else:
                probs = logits.softmax(dim=-1)
            picked = probs.multinomial(num_samples=1).view(*logits.shape[:-1], 1)
            ids = picked if ids is None else ids.gather(dim=-1, index=picked)

        ids = ids.flatten()
        self.output_ids[self.live_sample_mask, self.current_index] = ids.type_as(self.output_ids)
        self.current_index += 1
        continuing = (ids != self.end_token) & (self.current_index < self.output_ids.shape[-1])
        self.live_sample_mask[self.live_sample_mask.clone()] &= continuing
        return continuing

    def should_stop(self) -> bool:
        return self.current_index >= self.output_ids.shape[-1] - 1 or not self.live_sample_mask.any()

    def final_output(self) -> Tensor:
        return self.output_ids[:, 1:]
