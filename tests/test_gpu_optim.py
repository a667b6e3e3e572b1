"""Fused multi-tensor gradient clipping + RAdam (csrc/optim.cu) against the reference's arithmetic restated in
float64 on the CPU (sparse_vae/core/rectified_adam.py:15-88, core/language_model.py:120-122)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle_radam(params, grads, m, v, lr, beta1, beta2, eps, wd, step):
    """Literal float64 restatement of one reference RAdam.step (lamb=False) for a list of tensors."""
    beta2_t = beta2 ** step
    bias_v = (1 - beta2_t) ** 0.5
    rho_inf = 2.0 / (1.0 - beta2) - 1.0
    rho_t = rho_inf - 2 * step * beta2_t / (1 - beta2_t)
    if rho_t > 4:
        r_t = (((rho_t - 4.0) * (rho_t - 2.0) * rho_inf) / ((rho_inf - 4.0) * (rho_inf - 2.0) * rho_t)) ** 0.5
        lr = lr * r_t * bias_v
    bias_m = 1 - beta1 ** step
    for p, g, mm, vv in zip(params, grads, m, v):
        mm.mul_(beta1).add_(g, alpha=1 - beta1)
        vv.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        p.mul_(1 - lr * wd)
        if rho_t > 4:
            denom = (vv.sqrt() / bias_v).add_(eps)
            p.addcdiv_(mm, denom, value=-lr / bias_m)
        else:
            p.add_(mm, alpha=-lr / bias_m)


SHAPES = [(512, 512), (32768, 512), (2048,), (7,), (3, 5, 11), (65536 + 3,), (1,)]


@pytest.mark.parametrize('steps', [1, 4, 8])
def test_radam_matches_reference_rule(steps):
    from sparse_vae_b200.core.rectified_adam import RAdam
    g = torch.Generator().manual_seed(7295)
    params = [torch.randn(s, generator=g) * 0.02 for s in SHAPES]
    ref_p = [p.double().clone() for p in params]
    ref_m = [torch.zeros_like(p) for p in ref_p]
    ref_v = [torch.zeros_like(p) for p in ref_p]
    dev_p = [torch.nn.Parameter(p.cuda()) for p in params]
    opt = RAdam(dev_p, lr=3e-4, weight_decay=0.01)
    for step in range(1, steps + 1):          # steps 1..5 are the un-rectified (rho_t <= 4) branch
        grads = [torch.randn(s, generator=g) * (0.1 + 0.05 * step) for s in SHAPES]
        for p, gr in zip(dev_p, grads):
            p.grad = gr.cuda()
        opt.step()
        _oracle_radam(ref_p, [gr.double() for gr in grads], ref_m, ref_v, 3e-4, 0.9, 0.999, 1e-6, 0.01, step)
    for p, r in zip(dev_p, ref_p):
        torch.testing.assert_close(p.detach().cpu().double(), r, rtol=1e-5, atol=1e-7)
    for p, rm, rv in zip(dev_p, ref_m, ref_v):
        torch.testing.assert_close(opt.state[p]['exp_avg'].cpu().double(), rm, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(opt.state[p]['exp_avg_sq'].cpu().double(), rv, rtol=1e-5, atol=1e-11)
    assert opt.param_groups[0]['step'] == steps + 1
    assert '_fused_step' not in opt.state_dict()['param_groups'][0]


@pytest.mark.parametrize('max_norm', [1e9, 5.0, 0.01])
def test_clip_grad_norm_matches_torch(max_norm):
    from sparse_vae_b200.fused_optim import FusedGradClipper
    g = torch.Generator().manual_seed(1)
    grads = [torch.randn(s, generator=g) for s in SHAPES]
    # non-16-byte-aligned views exercise the scalar path
    flat = torch.randn(1000 + 1, generator=g).cuda()
    dev = [x.cuda() for x in grads] + [flat[1:]]
    ref = [x.double() for x in grads] + [flat[1:].cpu().double()]
    norm = math.sqrt(sum(float((x * x).sum()) for x in ref))
    coef = min(1.0, max_norm / (norm + 1e-6))
    clipper = FusedGradClipper()
    out = clipper(dev, max_norm)
    assert out.ndim == 0 and out.is_cuda
    assert abs(out.item() - norm) <= 2e-6 * norm
    for d, r in zip(dev, ref):
        torch.testing.assert_close(d.cpu().double(), r * coef, rtol=3e-6, atol=1e-12)
    # deterministic: a second clipper on identical input gives the bit-identical norm
    dev2 = [x.cuda() for x in grads] + [flat[1:].clone()]
    for d, r in zip(dev2, ref):
        d.copy_(r.float())
    assert FusedGradClipper()(dev2, 1e9).item() == FusedGradClipper()([x.clone() for x in dev2], 1e9).item()


def test_fused_optim_rejects_cpu_tensors():
    from sparse_vae_b200.fused_optim import FusedGradClipper
    with pytest.raises(ValueError):
        FusedGradClipper()([torch.zeros(4)], 1.0)
