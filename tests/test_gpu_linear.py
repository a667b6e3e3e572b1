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


def test_weight_shadows_give_identical_training_steps():
    """One multi-tensor cast per step instead of one autocast cast per weight: same rounding of the same fp32 weights,
    so the first loss is bit-identical; gradients (and the losses after an update) agree up to the run-to-run noise
    of the backward pass (fp32 atomics for the global key block, whose results are then rounded to bf16: single
    elements move by a bf16 ulp between two runs of the SAME configuration)."""
    import sparse_vae_b200 as sv
    from sparse_vae_b200.core.lightning_shim import to_attrdict
    from sparse_vae_b200.core.linear import WeightShadows
    from sparse_vae_b200.synthetic import synthetic_tokens, to_device
    dev = torch.device('cuda')

    def run(enabled):
        WeightShadows.ENABLED = enabled
        try:
            torch.manual_seed(3)
            model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams(d_model=256, num_layers=4, latent_depth=16))).to(dev).eval()
            model.initialize_weights()
            (opt,), _ = model.configure_optimizers(tokens_per_batch=4 * 2048)
            batch = to_device(synthetic_tokens(4, 2048, seed=1), dev)
            losses, first_grads = [], None
            for step in range(3):
                for p in model.parameters():
                    p.grad = None
                torch.manual_seed(100 + step)
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    loss = model.training_step(batch, step)['loss']
                loss.backward()
                losses.append(loss.detach().clone())
                if step == 0:                     # before any update: only the atomics' summation order differs
                    first_grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
                opt.step()
            used = model.__dict__.get('_weight_shadows')
            return losses, first_grads, used
        finally:
            WeightShadows.ENABLED = True

    la, ga, used = run(True)
    lb, gb, unused = run(False)
    assert used is not None and used.epoch == 3 and len(used._params) > 40 and unused is None
    assert torch.equal(la[0], lb[0]), (la, lb)          # forward only: deterministic, so bit-identical
    assert all(abs(a.item() - b.item()) <= 1e-5 * abs(b.item()) for a, b in zip(la, lb)), (la, lb)
    assert ga.keys() == gb.keys()
    for k in ga:
        assert (ga[k] - gb[k]).abs().max().item() <= 2e-2 * gb[k].abs().max().item() + 1e-9, (k, (ga[k] - gb[k]).abs().max().item(), gb[k].abs().max().item())


def test_weight_shadows_refuse_a_backward_after_the_weights_changed():
    from sparse_vae_b200.core.linear import Linear, WeightShadows
    dev = torch.device('cuda')
    lin = Linear(512, 512).to(dev)
    shadows = WeightShadows(lin)
    x = torch.randn(2048, 512, device=dev, requires_grad=True)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        with shadows.step():
            y = lin(x)
        with torch.no_grad():
            lin.weight.mul_(2.0)                    # an optimizer step between forward and backward ...
        with shadows.step():                        # ... and another forward that refreshes the copies
            lin(x)
    with pytest.raises(RuntimeError, match='modified'):
        y.sum().backward()


def test_linear3_matches_three_separate_projections():
    """q / k / v projections of one input: outputs bit-identical, gradients equal up to the rounding of the summed
    input gradient (GEMM-epilogue accumulation vs two bf16 adds)."""
    from sparse_vae_b200.core.linear import Linear, _Linear3Fn, linear3
    dev = torch.device('cuda')
    torch.manual_seed(0)
    lins = [Linear(512, 512).to(dev) for _ in range(3)]
    x = torch.randn(4, 1024, 512, device=dev, requires_grad=True)
    gs = [torch.randn(4, 1024, 512, device=dev) for _ in range(3)]

    def run(fn):
        x.grad = None
        for m in lins:
            m.zero_grad()
        with torch.autocast('cuda', dtype=torch.bfloat16):
            outs = fn()
        torch.autograd.backward(outs, [g.to(o.dtype) for g, o in zip(gs, outs)])
        return [o.detach().clone() for o in outs], x.grad.clone(), [(m.weight.grad.clone(), m.bias.grad.clone()) for m in lins]

    o1, dx1, p1 = run(lambda: linear3(x, *lins))
    assert isinstance(linear3(x.detach().requires_grad_(), *lins)[0].grad_fn, type(None)) is False
    o2, dx2, p2 = run(lambda: tuple(m(x) for m in lins))
    assert all(torch.equal(a, b) for a, b in zip(o1, o2))
    assert (dx1 - dx2).abs().max().item() <= 2e-2 * dx2.abs().max().item()
    for (w1, b1), (w2, b2) in zip(p1, p2):
        assert torch.equal(w1, w2) and torch.equal(b1, b2)


@pytest.mark.parametrize('rows,n', [(65536, 512), (777, 2048), (5, 8)])
def test_colsum_single_launch_equals_two_launch(rows, n):
    """The last-block reduction sums the slabs in the same fixed order as the separate second launch: bit-identical,
    and the ticket counters are back at zero afterwards (reusable)."""
    from sparse_vae_b200 import _native as N
    dev = torch.device('cuda')
    x = torch.randn(rows, n, device=dev).to(torch.bfloat16)
    ws_floats = N.lib.svae_colsum_workspace_floats(rows, n)
    ws = torch.empty(ws_floats, device=dev)
    counters = torch.zeros(N.lib.svae_colsum_counters(n), device=dev, dtype=torch.int32)
    outs = []
    for ctr in (None, counters, counters):
        out = torch.empty(n, device=dev)
        N.check(N.lib.svae_colsum(x.data_ptr(), N.svae_dtype(x.dtype), rows, n, n, out.data_ptr(), ws.data_ptr(), ws_floats,
                                  None if ctr is None else ctr.data_ptr(), N.current_stream(dev)), 'svae_colsum')
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]) and int(counters.abs().sum()) == 0
    assert (outs[0].double() - x.double().sum(0)).abs().max().item() <= 1e-3 * max(1.0, x.double().sum(0).abs().max().item())


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
def test_residual_dropout_add_mask_scaling_and_gradients(dtype):
    """x + dropout(h): kept elements carry dtype(h / (1-p)), the rest nothing; the backward pass regenerates the same
    mask; keep rate and independence from torch's own dropout stream are statistical."""
    from sparse_vae_b200.core.residual import _ResidualDropoutAddFn, residual_dropout_add
    dev = torch.device('cuda')
    torch.manual_seed(1)
    drop = torch.nn.Dropout(p=0.1).train()
    shape = (8, 4096, 512)
    x = torch.randn(*shape, device=dev, requires_grad=True)
    h = (torch.randn(*shape, device=dev) + 3.0).to(dtype).requires_grad_()        # no zeros: the mask is visible in out - x
    before = torch.cuda.default_generators[0].get_offset()
    out = residual_dropout_add(x, h, drop)
    assert isinstance(out.grad_fn, _ResidualDropoutAddFn._backward_cls)
    assert torch.cuda.default_generators[0].get_offset() == before + 4
    branch = (out - x).detach()
    kept = branch != 0
    rate = kept.float().mean().item()
    assert abs(rate - 0.9) < 2e-3, rate
    want = (h.detach().float() * (1.0 / (1.0 - 0.1))).to(dtype).float()
    # out = fp32(x + branch): compare the branch through the same addition
    assert torch.equal(out.detach()[kept], (x.detach() + want)[kept]) and torch.equal(out.detach()[~kept], x.detach()[~kept])
    g = torch.randn(*shape, device=dev)
    out.backward(g)
    assert torch.equal(x.grad, g)
    want_dh = (g.to(dtype).float() * (1.0 / (1.0 - 0.1))).to(dtype)
    assert torch.equal(h.grad[kept], want_dh[kept]) and int(h.grad[~kept].abs().sum()) == 0
    # a second call draws a different mask; neighbouring elements are not correlated
    out2 = residual_dropout_add(x, h, drop)
    kept2 = (out2 - x).detach() != 0
    both = (kept & kept2).float().mean().item()
    assert abs(both - 0.81) < 3e-3, both
    pair = (kept[..., 1:] & kept[..., :-1]).float().mean().item()
    assert abs(pair - 0.81) < 3e-3, pair
    # eval mode: the plain residual add
    assert torch.equal(residual_dropout_add(x, h, drop.eval()), x + h)
