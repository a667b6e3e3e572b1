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
        """Turns the live samples' next-token logits into tokens, appends them and returns which samples go on."""
        if self.repetition_penalty > 1.0:
            self._penalise_repeats(logits)
        greedy = self.temperature <= 0.0 or self.top_k == 1
        tokens = (logits.argmax(dim=-1) if greedy else self._draw(logits / self.temperature)).flatten()

        self.output_ids[self.live_sample_mask, self.current_index] = tokens.type_as(self.output_ids)
        self.current_index += 1
        continuing = (tokens != self.end_token) & (self.current_index < self.output_ids.shape[-1])
        self.live_sample_mask[self.live_sample_mask.clone()] &= continuing
        return continuing

    def _penalise_repeats(self, logits: Tensor):
        """Tokens generated in the last 512 steps become less likely: positive logits divided, negative multiplied."""
        first = max(self.current_index - 512, 0)
        recent = self.output_ids[self.live_sample_mask, first:self.current_index]
        values = logits.gather(dim=-1, index=recent)
        values = torch.where(values < 0.0, values * self.repetition_penalty, values / self.repetition_penalty)
        logits.scatter_(dim=-1, index=recent, src=values)

    def _draw(self, logits: Tensor) -> Tensor:
        """top-k, then nucleus filtering on the sorted distribution, then one multinomial draw per sample."""
        candidates = None                                   # token ids of the columns of `logits`, None = identity
        if self.top_k > 0:
            logits, candidates = logits.topk(k=max(self.top_k, 1), sorted=False)
        if self.top_p < 1.0:
            logits, order = logits.sort(descending=True)
            candidates = order if candidates is None else candidates.gather(dim=-1, index=order)
            probs = logits.softmax(dim=-1)
            beyond = probs.cumsum(dim=-1) > self.top_p      # cumulative mass already past top_p: dropped ...
            beyond[..., :1] = False                         # ... except the most likely token
            probs[beyond] = 0.0
        else:
            probs = logits.softmax(dim=-1)
        column = probs.multinomial(num_samples=1).view(*logits.shape[:-1], 1)
        return column if candidates is None else candidates.gather(dim=-1, index=column)

    def should_stop(self) -> bool:
        return self.current_index >= self.output_ids.shape[-1] - 1 or not self.live_sample_mask.any()

    def final_output(self) -> Tensor:
        return self.output_ids[:, 1:]
