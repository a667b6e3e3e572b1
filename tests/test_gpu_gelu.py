"""erf-GELU kernels (csrc/gelu.cu, core/gelu.py) against float64 torch and against the reference's op sequence
`nn.Sequential(nn.Linear, nn.GELU(), nn.Linear)` (core/transformer_layer.py:20-24)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _exact(x64):
    cdf = torch.special.ndtr(x64)
    return x64 * cdf, cdf + x64 * torch.exp(-0.5 * x64 * x64) / math.sqrt(2 * math.pi)


def _ulp(dtype):
    return 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11          # half a unit in the last place, relative


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
@pytest.mark.parametrize('rows,n', [(1, 8), (3, 24), (1000, 520), (4096, 2048), (65536, 512)])
def test_gelu_forward_backward_match_float64(dtype, rows, n):
    from sparse_vae_b200.core.gelu import gelu_forward, gelu_backward
    g = torch.Generator().manual_seed(rows + n)
    x = (torch.randn(rows, n, generator=g) * 2.0).to('cuda', dtype)
    x.view(-1)[:8] = torch.tensor([0.0, -0.0, 6.0, -6.0, 9.5, -9.5, 1e-3, -1e-3], dtype=dtype, device='cuda')
    dy = torch.randn(rows, n, generator=g).to('cuda', dtype)
    y64, d64 = _exact(x.double())
    y = gelu_forward(x)
    # within half a 16-bit ulp of the exact value plus the 2e-6 (absolute, on Phi) of the polynomial
    assert ((y.double() - y64).abs() <= _ulp(dtype) * y64.abs() + 2.5e-6 * x.double().abs() + 1e-7).all()
    dx, sums = gelu_backward(dy, x, want_colsum=True)
    want = dy.double() * d64
    assert ((dx.double() - want).abs() <= _ulp(dtype) * want.abs() + 4e-6 * dy.double().abs() + 1e-7).all()
    # the column sums are those of the fp32 products (before the 16-bit rounding), accumulated in fp32
    assert (sums.double() - want.sum(0)).abs().max() <= 1e-5 * want.abs().sum(0).max() + 1e-6
    # deterministic, and the in-place form gives the same bits
    dx2, sums2 = gelu_backward(dy.clone(), x, want_colsum=True, inplace=True)
    assert torch.equal(dx, dx2) and torch.equal(sums, sums2)
    assert torch.equal(gelu_backward(dy, x)[0], dx)


def test_gelu_close_to_aten_rounding():
    from sparse_vae_b200.core.gelu import gelu_forward
    x = (torch.randn(1 << 22, generator=torch.Generator().manual_seed(1)) * 1.5).to('cuda', torch.bfloat16).view(-1, 2048)
    mine, aten = gelu_forward(x), torch.nn.functional.gelu(x)
    differ = (mine != aten)
    assert differ.float().mean().item() < 0.01          # ~0.4 %: one bf16 ulp where the fp32 value sits next to a rounding boundary
    big = x.float() > -3.0                                # (torch's 1 + erff(x / sqrt 2) loses its digits in the negative tail)
    assert ((mine.float() - aten.float()).abs()[big] <= 2.0 ** -7 * aten.float().abs()[big]).all()


def test_gelu_module_autograd():
    from sparse_vae_b200.core.gelu import GELU
    x = torch.randn(64, 1024, device='cuda', dtype=torch.bfloat16)
    a, b = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    dy = torch.randn_like(x)
    GELU()(a).backward(dy)
    torch.nn.GELU()(b).backward(dy)
    assert (a.grad.float() - b.grad.float()).abs().max() <= 2.0 ** -7 * b.grad.float().abs().max()
    assert GELU()(x.float()).dtype == torch.float32       # fp32 tensors take ATen's kernel
    assert set(GELU().state_dict()) == set()


@pytest.mark.parametrize('bias2', [False, True])
def test_fused_ffn_matches_reference_sequence(bias2):
    from sparse_vae_b200.core.gelu import GELU, ffn_forward, _FfnFn
    from sparse_vae_b200.core.linear import Linear
    torch.manual_seed(5)
    d = 512
    ours = torch.nn.Sequential(Linear(d, 4 * d), GELU(), Linear(4 * d, d, bias=bias2)).cuda()
    ref = torch.nn.Sequential(torch.nn.Linear(d, 4 * d), torch.nn.GELU(), torch.nn.Linear(4 * d, d, bias=bias2)).cuda()
    ref.load_state_dict(ours.state_dict())
    x = torch.randn(2, 2048, d, device='cuda')
    dy = torch.randn(2, 2048, d, device='cuda').to(torch.bfloat16)
    a, b = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        ya = ffn_forward(ours, a)
        yb = ref(b)
    assert ya.grad_fn is not None and type(ya.grad_fn).__name__ == '_FfnFnBackward'
    assert ya.dtype == torch.bfloat16

    def close(u, v, tol):
        return (u.float() - v.float()).abs().max().item() <= tol * v.float().abs().max().item()

    assert close(ya, yb, 1e-2)
    ya.backward(dy)
    yb.backward(dy)
    assert close(a.grad, b.grad, 1e-2)
    for (n1, p1), (n2, p2) in zip(ours.named_parameters(), ref.named_parameters()):
        assert n1 == n2 and p1.grad.dtype == torch.float32
        assert close(p1.grad, p2.grad, 1e-2), n1
    # fp64 reference of the same computation: the fused path is at least as close as the library sequence
    w1, b1, w2 = (t.detach().double() for t in (ours[0].weight, ours[0].bias, ours[2].weight))
    x64 = x.double().requires_grad_(True)
    h = torch.nn.functional.linear(x64, w1, b1)
    y64 = torch.nn.functional.linear(h * torch.special.ndtr(h), w2, ours[2].bias.detach().double() if bias2 else None)
    y64.backward(dy.double())
    assert close(a.grad, x64.grad, 2e-2)
