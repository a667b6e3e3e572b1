"""GPU: module-level parity -- the product TransformerVAE (fused kernels) against the golden fixture of the
reference's own training step (BASELINE config 1: B=2, L=512, d_model=256, 4 layers) and against the oracle."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).parent / 'golden'))
import make_golden as mg  # noqa: E402
from oracle import model as omodel  # noqa: E402

pytestmark = pytest.mark.gpu


def _build(case, dev):
    import sparse_vae_b200 as sv
    from sparse_vae_b200.core.lightning_shim import to_attrdict
    hp = to_attrdict(sv.TransformerVAEHparams(d_model=case['d_model'], num_layers=case['num_layers'],
                                              num_heads=case['num_heads'], attn_window_size=case['window'],
                                              latent_depth=case['latent']))
    model = sv.TransformerVAE(hp)
    weights = mg.model_params([(n, tuple(p.shape)) for n, p in model.named_parameters()])
    with torch.no_grad():
        for n, p in model.named_parameters():
            p.copy_(torch.tensor(weights[n]))
    return sv, model.to(dev).eval()          # eval(): Dropout(0.1) off, as in the fixture


def _batch(sv, case, dev):
    tok = torch.tensor(mg.model_tokens(case), device=dev)
    lengths = torch.tensor(case['lengths'], device=dev)
    return {'token_ids': sv.PaddedTensor.from_raw(tok), 'num_tokens': lengths, 'num_bytes': 4 * lengths}


def test_fp32_training_step_matches_reference_fixture(golden_dir, monkeypatch):
    g = np.load(golden_dir / 'model_golden.npz')
    case = mg.MODEL_CASE
    dev = torch.device('cuda')
    sv, model = _build(case, dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    # eps of the fixture came from torch's CPU generator; feed the same eps by choosing mu/sigma-independent
    # injection: run sample_z, then replace z with loc + eps*scale computed from the fused sigma.
    eps = torch.tensor(g['eps'], device=dev)
    orig = model.sample_z

    def sample_z_with_fixture_eps(encoder_out, token_counts, stage='train'):
        z, kl, q = orig(encoder_out, token_counts, stage)
        return q.loc + eps * q.scale, kl, q

    monkeypatch.setattr(model, 'sample_z', sample_z_with_fixture_eps)
    out = model.training_step(_batch(sv, case, dev), 0)
    loss = out['loss']
    assert abs(loss.item() - float(g['loss'])) <= 1e-4 * abs(float(g['loss']))
    assert abs(model.logged['train_nll'].item() - float(g['nll'])) <= 1e-4 * abs(float(g['nll']))
    assert abs(model.logged['train_kl'].item() - float(g['raw_kl_mean'])) <= 1e-5 * abs(float(g['raw_kl_mean']))
    assert torch.allclose(out['posterior'].loc.cpu(), torch.tensor(g['posterior_loc']), rtol=1e-3, atol=1e-5)
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters()}
    assert sorted(n for n, gr in grads.items() if gr is None) == list(g['no_grad'])
    gn = torch.sqrt(sum((gr.double() ** 2).sum() for gr in grads.values() if gr is not None)).item()
    assert abs(gn - float(g['grad_norm'])) <= 1e-3 * float(g['grad_norm'])
    for key in g.files:
        if key.startswith('grad.'):
            ref = torch.tensor(g[key])
            got = grads[key[5:]].cpu()
            assert (got - ref).abs().max().item() <= 5e-3 * ref.abs().max().item() + 1e-7, key


def test_bf16_autocast_step_close_to_oracle(golden_dir):
    g = np.load(golden_dir / 'model_golden.npz')
    case = mg.MODEL_CASE
    dev = torch.device('cuda')
    sv, model = _build(case, dev)
    torch.manual_seed(7295)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        out = model.training_step(_batch(sv, case, dev), 0)
    out['loss'].backward()
    # nll is insensitive to the particular eps; the golden fp32 value is the yardstick
    assert abs(model.logged['train_nll'].item() - float(g['nll'])) <= 2e-2 * abs(float(g['nll']))
    assert abs(model.logged['train_kl'].item() - float(g['raw_kl_mean'])) <= 2e-2 * abs(float(g['raw_kl_mean']))
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)


def test_grad_checkpointing_recompute_is_deterministic():
    case = dict(mg.MODEL_CASE)
    dev = torch.device('cuda')
    sv, model = _build(case, dev)
    batch = _batch(sv, case, dev)
    losses, norms = [], []
    for ckpt in (False, True):
        model.hparams.grad_checkpointing = ckpt
        model.zero_grad(set_to_none=True)
        torch.manual_seed(1)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            loss = model.training_step(batch, 0)['loss']
        loss.backward()
        losses.append(loss.item())
        norms.append(model.decoder_layers[0].attention.q_linear.weight.grad.float().norm().item())
    assert losses[0] == losses[1]
    assert abs(norms[0] - norms[1]) <= 1e-3 * norms[0]


def test_fused_head_matches_unfused_logits_path():
    """`training_step` (fused vocabulary head + cross-entropy) against `reconstruct` -> `get_nll` (materialised logits,
    the reference's literal path) on the same model, same z: identical objective."""
    case = mg.MODEL_CASE
    dev = torch.device('cuda')
    sv, model = _build(case, dev)
    batch = _batch(sv, case, dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    from sparse_vae_b200.core.padded_tensor import split_padding
    tokens, padding = split_padding(batch['token_ids'])
    original = tokens.long()
    padding = original.eq(0) if padding is None else padding
    with torch.no_grad():
        x = model.input_layer(original)
        z = torch.randn(original.shape[0], 1, case['latent'], device=dev)
        logits = model.reconstruct(x, z, padding=padding)[..., :-1, :]
        nll_ref = model.get_nll(logits, original[..., 1:])
        hidden = model.reconstruct(x, z, padding=padding, return_hidden=True)
        from sparse_vae_b200.core.fused_ce import fused_vocab_nll
        nll = fused_vocab_nll(hidden, model.output_layer[-1], original[..., 1:])
    assert abs(nll.item() - nll_ref.item()) <= 1e-5 * abs(nll_ref.item())


def test_long_context_training_step_runs():
    """BASELINE config 4 shape per GPU: batch 4 x 16384 tokens, bf16; more than 2**30 logits -> two CE chunks."""
    import sparse_vae_b200 as sv
    from sparse_vae_b200.core.lightning_shim import to_attrdict
    from sparse_vae_b200.synthetic import synthetic_tokens, to_device
    dev = torch.device('cuda')
    torch.manual_seed(0)
    model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams())).to(dev)
    model.initialize_weights()
    batch = to_device(synthetic_tokens(4, 16384, seed=3), dev)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        loss = model.training_step(batch, 0)['loss']
    loss.backward()
    assert torch.isfinite(loss) and 9.5 < loss.item() < 11.5          # ~ln(32768) at random init
    gn = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in model.parameters() if p.grad is not None))
    assert torch.isfinite(gn) and gn.item() > 0


def test_autoregressive_sample_and_eval_paths_run():
    """`sample()` (KV-cache decoding, dense branch -- SURVEY 3.3) and `validation_step` on the GPU."""
    case = mg.MODEL_CASE
    dev = torch.device('cuda')
    sv, model = _build(case, dev)
    model.hparams.kl_weight = 1.0
    model.start_token, model.end_token = 1, 2
    with torch.no_grad():
        ids = model.sample(max_length=48, batch_size=3)
    assert ids.shape[0] == 3 and 1 <= ids.shape[1] <= 48 and ((ids >= 0) & (ids < 2 ** 15)).all()
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
        model.validation_step(_batch(sv, case, dev), 0)
    assert torch.isfinite(model.logged['val_nll']) and torch.isfinite(model.logged['val_loss'])


def test_generation_config_decoder_forward_in_chunks():
    """BASELINE config 5 shape: 256 samples x 4096 tokens on one GPU -- latents drawn from the prior, one teacher-forced
    decoder forward per chunk of 32 samples (the full logits would be 69 GB).  Size-independent property:
    samples are independent, so a sample decoded alone gives the logits it gets inside its chunk of 32."""
    import sparse_vae_b200 as sv
    from sparse_vae_b200.core.lightning_shim import to_attrdict
    from sparse_vae_b200.synthetic import synthetic_tokens, to_device
    dev = torch.device('cuda')
    torch.manual_seed(5)
    model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams())).to(dev).eval()
    model.initialize_weights()
    z_all = torch.randn(256, 1, model.hparams.latent_depth, device=dev)
    picked = None
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
        for chunk in range(8):
            tokens = to_device(synthetic_tokens(32, 4096, seed=100 + chunk), dev)['token_ids']
            x = model.input_layer(tokens.as_raw().long())
            logits = model.reconstruct(x, z_all[32 * chunk:32 * chunk + 32], padding=tokens.padding)
            assert logits.shape == (32, 4096, 32768) and torch.isfinite(logits[:, ::512]).all()
            if chunk == 3:
                picked = logits[17, :1024].float().clone()
                alone = model.reconstruct(x[17:18], z_all[32 * 3 + 17:32 * 3 + 18], padding=tokens.padding[17:18])
                alone = alone[0, :1024].float()
            del logits
    # bf16 GEMMs pick different tilings at batch 1 and 32; the attention kernel itself is batch-invariant
    assert (picked - alone).abs().max().item() <= 2e-2 * max(1.0, alone.abs().max().item())
