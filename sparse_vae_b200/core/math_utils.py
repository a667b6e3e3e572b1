"""Diagnostics used by the training step (reference surface: sparse_vae/core/math_utils.py:51-58)."""
import math

from torch.distributions import Normal


def marginal_kl(posteriors: Normal, num_samples: int = 10):
    """Monte-Carlo estimate of KL(q(z) || N(0, I)) for the aggregate posterior of a batch."""
    samples = posteriors.rsample([num_samples])
    cross = posteriors.log_prob(samples[:, :, None]).sum(dim=-1)
    log_q = cross.logsumexp(dim=2) - math.log(samples.shape[1])
    log_p = -0.5 * (samples.pow(2.0).sum(dim=-1).mean() + samples.shape[-1] * math.log(2 * math.pi))
    return log_p - log_q.mean()
