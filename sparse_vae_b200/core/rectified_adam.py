"""RAdam with optional LAMB trust ratio (reference surface: sparse_vae/core/rectified_adam.py).

Same hyper-parameters, state names (`exp_avg`, `exp_avg_sq`, per-group 1-indexed `step`) and update rule as the
reference.  CUDA fp32 parameter groups (lamb=False) take ONE fused multi-tensor kernel pass (csrc/optim.cu,
`svae_radam_step`); CPU tensors and LAMB groups are evaluated with torch._foreach ops.
"""
from __future__ import annotations

import torch
from torch.optim import Optimizer

from .. import _native as N


class RAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-6, lamb=False):
        assert 0.0 <= lr, "Learning rate must be non-negative"
        assert 0.0 <= eps, "Epsilon must be non-negative"
        assert 0.0 <= betas[0] < 1.0, "Adam beta1 must be between 0.0 and 1.0"
        assert 0.0 <= betas[1] < 1.0, "Adam beta2 must be between 0.0 and 1.0"
        assert 0.0 <= weight_decay, "Weight decay must be non-negative"
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, lamb=lamb))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()

        for group in self.param_groups:
            beta1, beta2 = group['betas']
            lr = group['lr']
            step = group.setdefault('step', 1)
            beta2_t = beta2 ** step
            bias_v = (1 - beta2_t) ** 0.5
            bias_m = 1 - beta1 ** step

            # variance rectification term of the adaptive learning rate
            rho_inf = 2.0 / (1.0 - beta2) - 1.0
            rho_t = rho_inf - 2 * step * beta2_t / (1 - beta2_t)
            rectified = rho_t > 4
            if rectified:
                r_t = (((rho_t - 4.0) * (rho_t - 2.0) * rho_inf) / ((rho_inf - 4.0) * (rho_inf - 2.0) * rho_t)) ** 0.5
                lr *= r_t * bias_v

            params, grads, m, v = [], [], [], []
            for p in group['params']:
                if p.grad is None:
                    continue
                assert not p.grad.is_sparse, 'RAdam does not support sparse gradients, use SparseAdam instead'
                state = self.state[p]
                if not state:
                    state['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                params.append(p); grads.append(p.grad); m.append(state['exp_avg']); v.append(state['exp_avg_sq'])
            if not params:
                group['step'] += 1
                continue

            if N.FUSED_EXTRAS and not group['lamb'] and all(
                    p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous()
                    and g.dtype == torch.float32 and a.is_contiguous() and b.is_contiguous()
                    for p, g, a, b in zip(params, grads, m, v)):
                from ..fused_optim import FusedRAdamStep
                cache = self.__dict__.setdefault('_fused_steps', {})     # kept off param_groups (state_dict stays clean)
                fused = cache.get(id(group))
                if fused is None:
                    fused = cache[id(group)] = FusedRAdamStep()
                from .graph_step import StepOptimArgs
                args_dev = StepOptimArgs.active.block_for(group) if StepOptimArgs.active is not None else 0
                fused(params, grads, m, v, group['lr'], beta1, beta2, group['eps'], group['weight_decay'], step, args_dev)
                group['step'] += 1
                continue

            torch._foreach_mul_(m, beta1)
            torch._foreach_add_(m, grads, alpha=1 - beta1)
            torch._foreach_mul_(v, beta2)
            torch._foreach_addcmul_(v, grads, grads, value=1 - beta2)

            def adam_direction():
                if rectified:
                    denom = torch._foreach_sqrt(v)
                    torch._foreach_div_(denom, bias_v)
                    torch._foreach_add_(denom, group['eps'])
                    return torch._foreach_div(m, denom)
                return [t.clone() for t in m]           # SGD with momentum while the variance is intractable

            if group['lamb']:
                updates = torch._foreach_mul(params, -group['weight_decay'])
                torch._foreach_add_(updates, adam_direction(), alpha=-1.0 / bias_m)
                for p, u in zip(params, updates):
                    trust = p.norm().clamp(min=0.01, max=10.0) / u.norm()
                    p.add_(u * (lr * trust))
                    self.state[p]['trust_ratio'] = trust
            else:
                torch._foreach_mul_(params, 1 - lr * group['weight_decay'])
                torch._foreach_add_(params, adam_direction(), alpha=-lr / bias_m)

            group['step'] += 1
        return loss
