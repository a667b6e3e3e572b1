"""Training-level equivalence of the widened rows: the same TransformerVAE, same weights, same batches and
seeds is trained for 30 steps (a) with the fused kernels and (b) with the reference's literal torch op sequence for
rotary / LayerNorm / Linear / vocabulary cross-entropy / clipping / RAdam (`_native.FUSED_EXTRAS = False`; the
attention and bottleneck kernels are the product path in both).  The loss curves must agree step by step within bf16
noise and the loss must go down by the same amount in both."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batches(n, B, L, dev):
    """Learnable synthetic data: arithmetic token progressions with a per-sequence stride."""
    g = torch.Generator().manual_seed(7295)
    out = []
    for _ in range(n):
        start = torch.randint(3, 2000, (B, 1), generator=g)
        stride = torch.randint(1, 4, (B, 1), generator=g)
        tok = 3 + (start + stride * torch.arange(L)[None, :]) % 2000
        tok[:, 0], tok[:, -1] = 1, 2
        counts = torch.full((B,), L)
        out.append({'token_ids': tok.to(dev), 'num_tokens': counts.to(dev), 'num_bytes': (4 * counts).to(dev)})
    return out


def _train(fused: bool, steps: int, dev):
    import sparse_vae_b200 as sv
    from sparse_vae_b200 import _native as N
    from sparse_vae_b200.core.lightning_shim import to_attrdict
    old = N.FUSED_EXTRAS
    N.FUSED_EXTRAS = fused
    try:
        torch.manual_seed(7295)
        hp = to_attrdict(sv.TransformerVAEHparams(d_model=256, num_layers=4, num_heads=4, lr=1e-3, grad_clip_threshold=5.0))
        model = sv.TransformerVAE(hp).to(dev)
        model.initialize_weights()
        model.train()
        for m in model.modules():                     # the fused x + dropout(h) draws its own Philox mask (tested in
            if isinstance(m, torch.nn.Dropout):       # test_gpu_linear.py); without dropout both runs see the same noise
                m.p = 0.0
        (opt,), _ = model.configure_optimizers(tokens_per_batch=100_000, accumulate_grad_batches=1)
        losses, norms = [], []
        torch.manual_seed(1234)                       # eps streams identical in both runs
        for batch in _batches(steps, 4, 512, dev):
            opt.zero_grad(set_to_none=True)
            with torch.autocast('cuda', dtype=torch.bfloat16):
                loss = model.training_step(batch, 0)['loss']
            loss.backward()
            model.on_after_backward()
            opt.step()
            model.global_step += 1
            losses.append(loss.item())
            norms.append(float(model.logged['grad_norm']))
        return losses, norms
    finally:
        N.FUSED_EXTRAS = old


def test_fused_rows_train_like_the_reference_op_sequence():
    dev = torch.device('cuda')
    steps = 30
    fused, fused_norms = _train(True, steps, dev)
    plain, plain_norms = _train(False, steps, dev)
    assert all(torch.isfinite(torch.tensor(fused))) and all(torch.isfinite(torch.tensor(plain)))
    # identical start (same weights, same batch, forward differs only by rounding)
    assert abs(fused[0] - plain[0]) <= 2e-3 * plain[0]
    assert abs(fused_norms[0] - plain_norms[0]) <= 2e-2 * plain_norms[0]
    # step-by-step agreement within bf16 training noise, and both learn
    rel = [abs(a - b) / b for a, b in zip(fused, plain)]
    assert max(rel[:10]) <= 2e-2, rel[:10]
    assert max(rel) <= 8e-2, rel
    # (RAdam's first steps are un-rectified momentum SGD: 30 steps only move the loss by ~0.1 nat, identically in both runs)
    assert fused[-1] < fused[0] - 0.05 and plain[-1] < plain[0] - 0.05, (fused[0], fused[-1], plain[0], plain[-1])
    assert abs((fused[0] - fused[-1]) - (plain[0] - plain[-1])) <= 0.1 * (plain[0] - plain[-1])


def test_training_with_fused_dropout_learns_like_torch_dropout():
    """With dropout on, the fused x + dropout(h) and nn.Dropout draw different masks, so the curves agree only
    statistically: same start within dropout noise, same amount learnt."""
    dev = torch.device('cuda')
    import sparse_vae_b200 as sv
    from sparse_vae_b200.core import residual
    from sparse_vae_b200.core.lightning_shim import to_attrdict

    def train(fused_dropout):
        original = residual.residual_dropout_add
        if not fused_dropout:
            residual_plain = lambda x, h, drop: residual.residual_add(x, drop(h))     # noqa: E731
            import sparse_vae_b200.core.attention as attn
            attn.residual_dropout_add = residual_plain
        try:
            torch.manual_seed(7295)
            hp = to_attrdict(sv.TransformerVAEHparams(d_model=256, num_layers=4, num_heads=4, lr=1e-3, grad_clip_threshold=5.0))
            model = sv.TransformerVAE(hp).to(dev)
            model.initialize_weights()
            model.train()
            (opt,), _ = model.configure_optimizers(tokens_per_batch=100_000, accumulate_grad_batches=1)
            torch.manual_seed(99)
            losses = []
            for batch in _batches(30, 4, 512, dev):
                opt.zero_grad(set_to_none=True)
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    loss = model.training_step(batch, 0)['loss']
                loss.backward()
                model.on_after_backward()
                opt.step()
                losses.append(loss.item())
            return losses
        finally:
            import sparse_vae_b200.core.attention as attn
            attn.residual_dropout_add = original

    a, b = train(True), train(False)
    assert abs(a[0] - b[0]) <= 1e-2 * b[0]
    assert a[-1] < a[0] - 0.05 and b[-1] < b[0] - 0.05
    assert abs((a[0] - a[-1]) - (b[0] - b[-1])) <= 0.25 * (b[0] - b[-1])
