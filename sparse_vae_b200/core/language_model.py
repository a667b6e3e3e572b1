"""`LanguageModel` base class and hyper-parameters (reference surface: sparse_vae/core/language_model.py).

Keeps the dataclass fields, the RAdam + cosine LambdaLR recipe with sqrt batch-size scaling, BERT-style
initialisation, `get_nll` with ignore_index 0 and the chunked cross-entropy, and gradient clipping in
`on_after_backward`.  Lightning-only pieces (callbacks, tokenizer hookup) are reduced to what runs without it.
"""
from __future__ import annotations

import math
from abc import ABC
from dataclasses import dataclass
from functools import partial
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn, Tensor
from torch.optim.lr_scheduler import LambdaLR

from .. import _native as N
from .lightning_shim import DictConfig, LightningModule
from .rectified_adam import RAdam


@dataclass
class LanguageModelHparams(ABC):
    grad_clip_threshold: float = 5.0
    init_scale: Optional[float] = 0.02      # stddev of the Gaussian init; None keeps PyTorch's defaults

    base_batch_size: int = 100_000          # reference batch size (tokens) of the sqrt learning-rate scaling
    lr: float = 2e-4
    lr_decay_steps: Optional[int] = 250_000

    start_token: Optional[int] = None
    end_token: Optional[int] = None

    early_stopping_metric: str = 'val_nll'
    log_samples: bool = True


def cdiv(a: int, b: int) -> int:
    return (a + b - 1) // b


class LanguageModel(LightningModule, ABC):
    def __init__(self, hparams: DictConfig):
        super().__init__()
        self.save_hyperparameters(hparams)
        self.example_input_array = None
        self.tokenizer = None
        self.start_token = self.hparams.get('start_token')
        self.end_token = self.hparams.get('end_token')

    def configure_optimizers(self, tokens_per_batch: Optional[int] = None, accumulate_grad_batches: int = 1):
        if tokens_per_batch is None:        # Lightning path: read it off the trainer like the reference
            tokens_per_batch = self.trainer.datamodule.hparams.tokens_per_batch
            accumulate_grad_batches = self.trainer.accumulate_grad_batches
        batch_size = tokens_per_batch * accumulate_grad_batches
        lr = self.hparams.lr * (batch_size / self.hparams.base_batch_size) ** 0.5
        opt = RAdam(self.parameters(), lr=lr, weight_decay=0.01)
        sched = LambdaLR(opt, partial(cosine_decay, self.hparams.lr_decay_steps))
        return [opt], [{'scheduler': sched, 'interval': 'step'}]

    def on_fit_start(self):
        self.initialize_weights()

    def initialize_weights(self):
        scale = self.hparams.init_scale
        if scale is None:
            return
        for module in self.modules():
            if isinstance(module, (nn.BatchNorm1d, nn.LayerNorm)):
                continue
            if isinstance(module, (nn.Embedding, nn.Linear)):
                module.weight.data.normal_(0.0, scale)
            bias = getattr(getattr(module, 'bias', None), 'data', None)
            if bias is not None:
                bias.zero_()

    def get_nll(self, logits: Tensor, labels: Tensor, stage: str = 'train', bytes_per_token: Tensor = None):
        if extra_dims := logits.ndim - labels.ndim - 1:
            labels = labels.expand(*logits.shape[:extra_dims], *labels.shape)
        nll = robust_cross_entropy(logits, labels)
        if stage == 'val' and bytes_per_token is not None and hasattr(self, 'token_weights'):
            nats_per_byte = robust_cross_entropy(logits, labels, weight=self.token_weights) * bytes_per_token
            self.log('val_bpb', nats_per_byte / math.log(2))
        self.log(stage + '_nll', nll)
        return nll

    def training_step(self, batch: Dict[str, Tensor], batch_index: int) -> Tensor:
        logits = self.forward(batch)[..., :-1, :]
        return self.get_nll(logits, batch['token_ids'][..., 1:].long())

    def on_after_backward(self):
        grads = [p.grad for p in self.parameters() if p.grad is not None]
        if N.FUSED_EXTRAS and grads and all(g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() for g in grads):
            # one fused multi-tensor pass (csrc/optim.cu) instead of a per-parameter norm + scale
            if getattr(self, '_fused_clipper', None) is None:
                from ..fused_optim import FusedGradClipper
                self._fused_clipper = FusedGradClipper()
            grad_norm = self._fused_clipper(grads, self.hparams.grad_clip_threshold)
        else:
            grad_norm = torch.nn.utils.clip_grad_norm_(self.parameters(), self.hparams.grad_clip_threshold)
        self.log('grad_norm', grad_norm, on_step=True)

    def validation_step(self, batch: Dict[str, Tensor], batch_index: int) -> Tensor:
        logits = self.forward(batch)[..., :-1, :]
        return self.get_nll(logits, batch['token_ids'][..., 1:].long(), stage='val')

    def test_step(self, batch: Dict[str, Tensor], batch_index: int):
        return self.validation_step(batch, batch_index)

    def sample(self, max_length: int, batch_size: int = 1, **kwargs):
        return None


def cosine_decay(decay_steps: int, cur_step: int):
    progress = cur_step / max(1, decay_steps)
    if progress >= 1.0:
        print("Learning rate decayed to 0.0. Halting training.")
        raise KeyboardInterrupt
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * progress)))


def cosine_decay_with_warmup(decay_steps: int, warmup_steps: int, cur_step: int):
    if cur_step < warmup_steps:
        return cur_step / warmup_steps
    if not decay_steps:
        return 1.0
    return cosine_decay(max(1, decay_steps - warmup_steps), cur_step - warmup_steps)


def get_cosine_decay_with_warmup_schedule(decay_steps: int, warmup_steps: int):
    return partial(cosine_decay_with_warmup, decay_steps, warmup_steps)


def robust_cross_entropy(logits, labels, weight=None):
    """Cross-entropy with padding id 0 ignored, evaluated in chunks of at most 2**30 logits along the sequence
    (reference core/language_model.py:161-170)."""
    chunks = cdiv(logits.numel(), 2 ** 30)
    if chunks == 1:
        return F.cross_entropy(logits.flatten(end_dim=1), labels.flatten(), ignore_index=0, weight=weight)
    return torch.stack([
        F.cross_entropy(lo.flatten(end_dim=1), la.flatten(), ignore_index=0, weight=weight)
        for lo, la in zip(logits.chunk(chunks, dim=-2), labels.chunk(chunks, dim=-1))
    ]).mean()
