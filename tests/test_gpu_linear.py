"""`Linear` with the column-sum bias-gradient kernel (csrc/colsum.cu, core/linear.py) against `nn.Linear`."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize('rows,n', [(1, 8), (63, 512), (4096, 512), (65536, 512), (5000, 2048), (777, 32768)])
def test_colsum_matches_float64(dtype, rows, n):
    from sparse_vae_b200.core.linear import colsum
    g = torch.Generator().manual_seed(rows + n)
    x = torch.randn(rows, n, generator=g).to('cuda', dtype)
    ref = x.double().sum(0)
    got = colsum(x)
    assert got.dtype == torch.float32
    assert (got.double() - ref).abs().max() <= 2e-6 * x.double().abs().sum(0).max() + 1e-6
    assert torch.equal(colsum(x), got)                       # deterministic


def test_colsum_strided_rows():
    from sparse_vae_b200.core.linear import colsum
    x = torch.randn(300, 1024, device='cuda', dtype=torch.bfloat16)
    view = x[:, 256:768]
    assert (colsum(view).double() - view.double().sum(0)).abs().max() <= 1e-3


@pytest.mark.parametrize('in_f,out_f', [(512, 512), (512, 2048)])
def test_linear_autocast_matches_nn_linear(in_f, out_f):
    from sparse_vae_b200.core.linear import Linear
    torch.manual_seed(3)
    ours, ref = Linear(in_f, out_f).cuda(), torch.nn.Linear(in_f, out_f).cuda()
    ref.load_state_dict(ours.state_dict())
    assert set(ours.state_dict()) == {'weight', 'bias'}
    x = torch.randn(4, 1024, in_f, device='cuda')
    dy = torch.randn(4, 1024, out_f, device='cuda').to(torch.bfloat16)
    a, b = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        ya, yb = ours(a), ref(b)
    assert ya.dtype == torch.bfloat16 and torch.equal(ya, yb)
    ya.backward(dy)
    yb.backward(dy)
    assert torch.equal(a.grad, b.grad)
    # reference: weight / bias gradients rounded to bf16 by the GEMM / reduction; ours: fp32 accumulation kept
    assert (ours.weight.grad - ref.weight.grad).abs().max() <= 8e-3 * ref.weight.grad.abs().max()
    assert (ours.bias.grad - ref.bias.grad).abs().max() <= 8e-3 * ref.bias.grad.abs().max()
    exact = dy.double().flatten(0, 1).sum(0)
    assert (ours.bias.grad.double() - exact).abs().max() <= (ref.bias.grad.double() - exact).abs().max() + 1e-6


def test_linear_plain_paths_are_nn_linear():
    from sparse_vae_b200.core.linear import Linear
    lin = Linear(64, 64)
    x = torch.randn(3, 64)
    torch.testing.assert_close(lin(x), torch.nn.functional.linear(x, lin.weight, lin.bias))       # CPU
    lin = lin.cuda()
    xc = torch.randn(2000, 64, device='cuda', requires_grad=True)
    torch.testing.assert_close(lin(xc), torch.nn.functional.linear(xc, lin.weight, lin.bias))     # no autocast


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
@pytest.mark.parametrize('shape', [(16, 4096, 512), (3, 1, 8), (2, 37, 256)])
def test_residual_add_equals_aten_mixed_add(dtype, shape):
    """fp32 stream + 16-bit branch: values and both gradients bit-identical to `x + h`."""
    from sparse_vae_b200.core.residual import _ResidualAddFn, residual_add
    dev = torch.device('cuda')
    torch.manual_seed(sum(shape))
    x = torch.randn(*shape, device=dev, requires_grad=True)
    h = torch.randn(*shape, device=dev).to(dtype).requires_grad_()
    g = torch.randn(*shape, device=dev)
    out = residual_add(x, h)
    assert isinstance(out.grad_fn, _ResidualAddFn._backward_cls)
    out.backward(g)
    got = (out.detach().clone(), x.grad.clone(), h.grad.clone())
    x.grad = h.grad = None
    ref = x + h
    ref.backward(g)
    assert torch.equal(got[0], ref) and torch.equal(got[1], x.grad) and torch.equal(got[2], h.grad)
    assert got[2].dtype == dtype
    # shapes the kernel does not take fall back to the plain expression
    y = residual_add(x[..., :7], h[..., :7])
    assert torch.equal(y, x[..., :7] + h[..., :7])
