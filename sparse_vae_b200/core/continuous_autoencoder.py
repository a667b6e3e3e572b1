"""`ContinuousVAE` (reference surface: sparse_vae/core/continuous_autoencoder.py).

`sample_z` is the bottleneck of the hot path: the reference's `q(encoder_out, get_kl=True)` -> `rsample()` ->
flatten/sum/div/mean (:42-52) is one fused launch here (ConditionalGaussian.sample), with the same return
values, the same `{stage}_kl` log (mean raw KL per sample) and the same Philox stream as `Normal.rsample`.
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod
from dataclasses import dataclass

import torch
from torch import Tensor
from torch.distributions import Normal

from .conditional_gaussian import ConditionalGaussian
from .language_model import LanguageModel, LanguageModelHparams


@dataclass
class ContinuousVAEHparams(LanguageModelHparams, ABC):
    latent_depth: int = 64

    kl_annealing_steps: int = 0
    kl_weight_start: float = 1.0
    kl_weight_end: float = 1.0
    kl_weight: float = 1.0

    early_stopping_metric: str = 'val_loss'


class ContinuousVAE(LanguageModel, ABC):
    q_of_z_given_x: ConditionalGaussian

    def on_train_start(self):
        self.hparams.kl_weight = self.hparams.kl_weight_start

    def on_after_backward(self):
        super().on_after_backward()
        hp = self.hparams
        if not hp.kl_annealing_steps or hp.kl_weight >= hp.kl_weight_end:
            return
        progress = self.global_step / hp.kl_annealing_steps
        hp.kl_weight = hp.kl_weight_start + (hp.kl_weight_end - hp.kl_weight_start) * progress

    def sample_z(self, encoder_out: Tensor, token_counts: Tensor, stage: str = 'train'):
        """Returns (z, kl, q_of_z): the latent sample, mean_b(KL_b / tokens_b) and the posterior."""
        z, kl, raw_kl, q_of_z = self.q_of_z_given_x.sample(encoder_out, token_counts)
        self.log(stage + '_kl', raw_kl.mean())
        return z, kl, q_of_z

    @staticmethod
    def prior_log_prob(z: Tensor):
        return -0.5 * z.pow(2.0).sum(dim=-1) - math.log(math.sqrt(2 * math.pi)) * z.shape[-1]

    def estimate_log_prob_iw(self, q_of_z: Normal, x: Tensor, labels: Tensor, num_samples: int, num_iter: int = 1,
                             padding: Tensor = None):
        """Importance-weighted estimate of log p(x) (evaluation only; reference :62-80)."""
        assert num_samples % num_iter == 0
        chunk = num_samples // num_iter
        log_ws = []
        for _ in range(num_iter):
            z = q_of_z.rsample([chunk])
            log_p_z = self.prior_log_prob(z)
            log_q_z = q_of_z.log_prob(z).sum(dim=-1)
            log_p_x = torch.stack([self.p_of_x_given_z(x, z[i], labels[..., 1:], padding=padding) for i in range(chunk)])
            log_ws.append(log_p_z.reshape(chunk, -1) + log_p_x - log_q_z.reshape(chunk, -1))
        return torch.cat(log_ws).logsumexp(dim=0) - math.log(num_samples)

    def p_of_x_given_z(self, x, z, labels, padding=None) -> Tensor:
        logits = self.reconstruct(x, z, padding=padding)[..., :-1, :]
        log_probs = logits.log_softmax(dim=-1)
        log_probs[..., 0] = 0.0     # padding tokens do not count
        return log_probs.gather(dim=-1, index=labels.unsqueeze(-1)).squeeze(-1).sum(dim=-1)

    @abstractmethod
    def reconstruct(self, x, z, padding=None) -> Tensor:
        raise NotImplementedError
