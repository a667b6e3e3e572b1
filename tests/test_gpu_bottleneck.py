"""GPU parity of the fused latent bottleneck (through the C ABI).

  * z is BIT-EXACT against torch's own `Normal(mu, sigma).rsample()` run with the same CUDA generator state
    (that is the reference's arithmetic: core/continuous_autoencoder.py:44), and the generator ends at the
    same Philox offset;
  * KL terms and gradients within 1e-5 relative of the oracle (oracle/bottleneck.py) and of the golden fixture
    produced by the reference's ConditionalGaussian + sample_z.
"""
import numpy as np
import pytest
import torch
from torch.distributions import Normal

from oracle import bottleneck as obn

pytestmark = pytest.mark.gpu


def _modules():
    import sparse_vae_b200 as sv
    from sparse_vae_b200.core import conditional_gaussian as cgm
    return sv, cgm


def _reference_sample(mulogvar, counts):
    """The reference's op sequence on the GPU (ConditionalGaussian.forward + sample_z) under bf16/fp32 promotion."""
    mu, logvar = mulogvar.chunk(2, dim=-1)
    var = logvar.float().exp()
    q = Normal(loc=mu, scale=var.sqrt(), validate_args=False)
    kl_elem = 0.5 * (mu.float() ** 2 + var - logvar.float() - 1.0)
    z = q.rsample()
    raw_kl = kl_elem.flatten(1).sum(dim=-1)
    kl = raw_kl.div(counts).mean()
    return z, q, kl_elem, raw_kl, kl


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32, torch.float16])
@pytest.mark.parametrize('B,latent', [(16, 64), (256, 64), (3, 48), (5, 37), (2, 1024)])
def test_z_bit_exact_with_torch_rsample(dtype, B, latent):
    sv, cgm = _modules()
    dev = torch.device('cuda')
    g = torch.Generator(device='cpu').manual_seed(B * 1000 + latent)
    mulogvar = (torch.randn(B, 1, 2 * latent, generator=g) * 0.7).to(dev, dtype)
    counts = torch.randint(32, 4096, (B,), generator=g).to(dev)

    torch.manual_seed(7295)
    z_ref, q_ref, kl_elem_ref, raw_ref, kl_ref = _reference_sample(mulogvar, counts)
    off_ref = torch.cuda.default_generators[0].get_offset()

    torch.manual_seed(7295)
    out = cgm.fused_bottleneck(mulogvar, counts)
    off = torch.cuda.default_generators[0].get_offset()

    assert off == off_ref
    assert out['z'].dtype == torch.float32 and out['z'].shape == (B, 1, latent)
    assert torch.equal(out['z'], z_ref.float())
    assert torch.equal(out['sigma'], q_ref.scale)
    assert torch.equal(out['kl_elem'], kl_elem_ref)
    assert torch.allclose(out['raw_kl'], raw_ref, rtol=1e-5, atol=1e-6)
    assert abs(out['kl'].item() - kl_ref.item()) <= 1e-5 * abs(kl_ref.item())


def test_large_stream_matches_torch_normal_and_oracle():
    """Beyond one sweep of the emulated ATen grid (numel > 256*1184) the kernel shares one Philox draw between
    four rows; the eps it implies must still be torch's."""
    sv, cgm = _modules()
    dev = torch.device('cuda')
    rows, latent = 20_000, 64                           # 1.28 M elements > 303,104
    mulogvar = torch.zeros(rows, 1, 2 * latent, device=dev)          # mu = 0, sigma = 1  ->  z == eps
    torch.manual_seed(123)
    eps_ref = torch.empty(rows, 1, latent, device=dev).normal_()
    off_ref = torch.cuda.default_generators[0].get_offset()
    torch.manual_seed(123)
    out = cgm.fused_bottleneck(mulogvar, torch.ones(rows, device=dev, dtype=torch.int64))
    assert torch.cuda.default_generators[0].get_offset() == off_ref
    assert torch.equal(out['z'], eps_ref)
    # CPU restatement of the stream (fp32; __sincosf vs libm differ in the last bits)
    want = obn.standard_normal_like_cuda(rows * latent, seed=123, offset=0,
                                         sm_count=torch.cuda.get_device_properties(0).multi_processor_count)
    assert np.abs(out['z'].flatten().cpu().numpy() - want).max() < 5e-6
    # non-shareable geometry (latent does not divide the grid) takes the per-element path
    rows2, latent2 = 9_000, 37
    ml2 = torch.zeros(rows2, 1, 2 * latent2, device=dev)
    torch.manual_seed(5)
    eps2 = torch.empty(rows2, 1, latent2, device=dev).normal_()
    torch.manual_seed(5)
    out2 = cgm.fused_bottleneck(ml2, torch.ones(rows2, device=dev, dtype=torch.int64))
    assert torch.equal(out2['z'], eps2)


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
def test_gradients_match_autograd_of_reference_ops(dtype):
    sv, cgm = _modules()
    dev = torch.device('cuda')
    B, latent = 16, 64
    g = torch.Generator(device='cpu').manual_seed(1)
    base = (torch.randn(B, 1, 2 * latent, generator=g) * 0.5).to(dev, dtype)
    counts = torch.randint(100, 4096, (B,), generator=g).to(dev)
    dz = torch.randn(B, 1, latent, generator=g).to(dev)
    w_kl = 0.37

    a = base.clone().requires_grad_(True)
    torch.manual_seed(99)
    z_ref, q_ref, _, raw_ref, kl_ref = _reference_sample(a, counts)
    ((z_ref.float() * dz).sum() + w_kl * kl_ref + 0.01 * raw_ref.sum() + 0.2 * q_ref.scale.sum()).backward()

    b = base.clone().requires_grad_(True)
    torch.manual_seed(99)
    out = cgm.fused_bottleneck(b, counts)
    ((out['z'] * dz).sum() + w_kl * out['kl'] + 0.01 * out['raw_kl'].sum() + 0.2 * out['sigma'].sum()).backward()

    tol = 1e-5 if dtype == torch.float32 else 1e-2      # bf16 gradients are rounded to bf16 on both sides
    denom = a.grad.float().abs().max().item()
    assert (a.grad.float() - b.grad.float()).abs().max().item() <= tol * denom


def test_matches_reference_golden_fixture(golden_dir):
    sv, cgm = _modules()
    dev = torch.device('cuda')
    g = np.load(golden_dir / 'bottleneck_golden.npz')
    cg = sv.ConditionalGaussian(g['enc'].shape[-1], g['W'].shape[0] // 2).to(dev)
    with torch.no_grad():
        cg.linear.weight.copy_(torch.tensor(g['W']))
        cg.linear.bias.copy_(torch.tensor(g['bias']))
    x = torch.tensor(g['enc'], device=dev)
    q, kl_elem = cg(x, get_kl=True)
    assert torch.allclose(kl_elem.cpu(), torch.tensor(g['kl_elem']), rtol=1e-4, atol=1e-6)
    assert torch.allclose(q.scale.cpu(), torch.tensor(g['sigma']), rtol=1e-5)
    assert torch.allclose(q.loc.cpu(), torch.tensor(g['loc']), rtol=1e-4, atol=1e-6)
    z, kl, raw_kl, _ = cg.sample(x, torch.tensor(g['counts'], device=dev))
    assert abs(kl.item() - float(g['kl'])) <= 1e-5 * abs(float(g['kl'])) + 1e-7
    assert abs(raw_kl.mean().item() - float(g['raw_kl'])) <= 1e-5 * abs(float(g['raw_kl']))
    # eps implied by the fused z must be a standard normal draw; the oracle reproduces z from it
    eps = (z - q.loc) / q.scale
    f = obn.bottleneck_forward(cg.linear(x).detach().cpu(), torch.tensor(g['counts']), eps.detach().cpu())
    assert torch.allclose(f['z'], z.detach().cpu(), rtol=1e-5, atol=1e-6)


def test_sample_z_logs_and_shapes():
    sv, _ = _modules()
    from sparse_vae_b200.core.lightning_shim import to_attrdict
    dev = torch.device('cuda')
    model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams(d_model=128, num_layers=4, num_heads=2))).to(dev)
    enc = torch.randn(4, 1, 128, device=dev)
    z, kl, q = model.sample_z(enc, torch.tensor([100, 200, 300, 400], device=dev))
    assert z.shape == (4, 1, 64) and kl.ndim == 0 and q.loc.shape == (4, 1, 64)
    assert 'train_kl' in model.logged
