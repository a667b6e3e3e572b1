"""CPU: host-side logic of the SURVEY 8f rows -- the per-token weights that reproduce `robust_cross_entropy`'s chunked
mean (reference core/language_model.py:161-170), the RAdam rule on CPU tensors (reference core/rectified_adam.py:15-88),
the drop-in modules' CPU behaviour, and the autograd glue of the in-place latent-row replacement."""
import math

import pytest
import torch
import torch.nn.functional as F


def _reference_robust_ce(logits, labels, chunk_elems):
    """The reference function with its 2**30 constant made a parameter so that small tensors exercise the chunking."""
    chunks = -(-logits.numel() // chunk_elems)
    if chunks == 1:
        return F.cross_entropy(logits.flatten(end_dim=1), labels.flatten(), ignore_index=0)
    return torch.stack([F.cross_entropy(lo.flatten(end_dim=1), la.flatten(), ignore_index=0)
                        for lo, la in zip(logits.chunk(chunks, dim=-2), labels.chunk(chunks, dim=-1))]).mean()


@pytest.mark.parametrize('B,S,V,chunk_elems', [(3, 17, 11, 2 ** 30), (3, 17, 11, 200), (2, 64, 8, 300), (4, 33, 5, 100),
                                               (1, 9, 7, 20)])
def test_token_weights_reproduce_chunked_mean(monkeypatch, B, S, V, chunk_elems):
    from sparse_vae_b200.core import fused_ce
    g = torch.Generator().manual_seed(B * S + V)
    logits = torch.randn(B, S, V, generator=g, dtype=torch.float64)
    labels = torch.randint(1, V, (B, S), generator=g)
    labels[0, S // 2:] = 0
    labels[-1, -1] = 0
    ref = _reference_robust_ce(logits, labels, chunk_elems)
    # _token_weights hard-codes 2**30 like the reference; scale the vocabulary argument so the same chunk count results
    chunks = -(-(B * S * V) // chunk_elems)
    fake_vocab = 1 if chunks == 1 else -(-((chunks - 1) * 2 ** 30 + 1) // (B * S))
    assert -(-(B * S * fake_vocab) // 2 ** 30) == chunks
    w = fused_ce._token_weights(labels, fake_vocab)
    nll = F.cross_entropy(logits.flatten(end_dim=1), labels.flatten(), ignore_index=0, reduction='none').view(B, S)
    got = (nll * w).sum()
    if torch.isnan(ref):                     # a chunk without any valid token: 0/0 in the reference, and here
        assert torch.isnan(got)
    else:
        assert abs(got.item() - ref.item()) <= 1e-6 * abs(ref.item())   # weights are fp32 (they scale an fp32 gradient)
    if not torch.isnan(ref):
        assert (w[labels == 0] == 0).all()


def test_radam_cpu_path_matches_reference_rule():
    from sparse_vae_b200.core.rectified_adam import RAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(5, 7), (11,), (2, 3, 4)]
    params = [torch.nn.Parameter(torch.randn(s, generator=g, dtype=torch.float64)) for s in shapes]
    ref_p = [p.detach().clone() for p in params]
    ref_m = [torch.zeros_like(p) for p in ref_p]
    ref_v = [torch.zeros_like(p) for p in ref_p]
    opt = RAdam(params, lr=1e-2, weight_decay=0.01)
    beta1, beta2, eps, wd = 0.9, 0.999, 1e-6, 0.01
    for step in range(1, 9):
        grads = [torch.randn(s, generator=g, dtype=torch.float64) for s in shapes]
        for p, gr in zip(params, grads):
            p.grad = gr.clone()
        opt.step()
        lr = 1e-2
        beta2_t = beta2 ** step
        bias_v = (1 - beta2_t) ** 0.5
        rho_inf = 2.0 / (1.0 - beta2) - 1.0
        rho_t = rho_inf - 2 * step * beta2_t / (1 - beta2_t)
        if rho_t > 4:
            lr *= (((rho_t - 4.0) * (rho_t - 2.0) * rho_inf) / ((rho_inf - 4.0) * (rho_inf - 2.0) * rho_t)) ** 0.5 * bias_v
        for p, gr, m, v in zip(ref_p, grads, ref_m, ref_v):
            m.mul_(beta1).add_(gr, alpha=1 - beta1)
            v.mul_(beta2).addcmul_(gr, gr, value=1 - beta2)
            p.mul_(1 - lr * wd)
            if rho_t > 4:
                p.addcdiv_(m, (v.sqrt() / bias_v).add_(eps), value=-lr / (1 - beta1 ** step))
            else:
                p.add_(m, alpha=-lr / (1 - beta1 ** step))
    for p, r in zip(params, ref_p):
        torch.testing.assert_close(p.detach(), r, rtol=1e-12, atol=1e-14)


def test_drop_in_modules_on_cpu_are_the_torch_modules():
    from sparse_vae_b200.core.layer_norm import LayerNorm
    from sparse_vae_b200.core.linear import Linear
    ln, lin = LayerNorm(128), Linear(128, 64)
    assert set(ln.state_dict()) == {'weight', 'bias'} and set(lin.state_dict()) == {'weight', 'bias'}
    assert isinstance(ln, torch.nn.LayerNorm) and isinstance(lin, torch.nn.Linear)
    x = torch.randn(3, 5, 128)
    torch.testing.assert_close(ln(x), F.layer_norm(x, (128,), ln.weight, ln.bias, ln.eps))
    torch.testing.assert_close(lin(x), F.linear(x, lin.weight, lin.bias))
    with torch.autocast('cpu', dtype=torch.bfloat16):
        assert lin(x).dtype == torch.bfloat16           # plain nn.Linear behaviour under CPU autocast


def test_round2_drop_in_modules_on_cpu_are_the_torch_modules():
    from sparse_vae_b200.core.embedding import Embedding
    from sparse_vae_b200.core.gelu import GELU, ffn_forward
    from sparse_vae_b200.core.linear import Linear
    gelu, emb = GELU(), Embedding(50, 8)
    assert isinstance(gelu, torch.nn.GELU) and isinstance(emb, torch.nn.Embedding)
    assert set(gelu.state_dict()) == set() and set(emb.state_dict()) == {'weight'}
    x = torch.randn(3, 5, 16)
    assert torch.equal(gelu(x), F.gelu(x))
    ids = torch.randint(0, 50, (4, 7))
    y = emb(ids)
    assert torch.equal(y, F.embedding(ids, emb.weight)) and type(y.grad_fn).__name__ == 'EmbeddingBackward0'
    ffn = torch.nn.Sequential(Linear(16, 64), GELU(), Linear(64, 16, bias=False))
    assert torch.equal(ffn_forward(ffn, x), ffn(x))             # CPU tensors: the plain Sequential


def test_row_stride_and_adjacency_helpers():
    from sparse_vae_b200.core.linear import _adjacent, _row_stride
    buf = torch.zeros(2, 6, 48, dtype=torch.bfloat16)
    q, k, v = buf[..., :16], buf[..., 16:32], buf[..., 32:]
    assert _row_stride(buf) == 48 and _row_stride(q) == 48 and _row_stride(k) == 48 and _row_stride(v) == 48
    assert _row_stride(buf.transpose(0, 1)) is None              # rows of different batches are not uniformly spaced
    assert _row_stride(buf[..., ::2]) is None                    # inner stride 2
    assert _row_stride(buf[..., 4:20]) is None                   # slice start not 16-byte aligned
    flat = torch.zeros(3 * 32 + 8, dtype=torch.bfloat16)
    a, b, c = (flat[i * 32:(i + 1) * 32].view(4, 8) for i in range(3))
    assert _adjacent([a, b, c], 32)
    assert not _adjacent([a, c, b], 32) and not _adjacent([a, b, None], 32)
    assert not _adjacent([a, b, torch.zeros(4, 8, dtype=torch.bfloat16)], 32)
    w_cat = torch.as_strided(a, (12, 8), (8, 1))
    assert w_cat.data_ptr() == a.data_ptr() and torch.equal(w_cat[4:8], b)


def test_fused_paths_refuse_cpu_tensors():
    from sparse_vae_b200.core.fused_ce import fused_vocab_nll, supported
    lin = torch.nn.Linear(8, 8192)
    assert not supported(torch.randn(1, 4, 8), lin)
    with pytest.raises(ValueError):
        fused_vocab_nll(torch.randn(1, 4, 8), lin, torch.ones(1, 3, dtype=torch.long))
    from sparse_vae_b200.fused_optim import FusedRAdamStep
    with pytest.raises(ValueError):
        FusedRAdamStep()([torch.zeros(4)], [torch.zeros(4)], [torch.zeros(4)], [torch.zeros(4)], 1e-3, 0.9, 0.999, 1e-6, 0.0, 1)


def test_replace_first_position_matches_cat_and_gradients():
    from sparse_vae_b200.transformer_vae import _ReplaceFirstPosition
    g = torch.Generator().manual_seed(3)
    base = torch.randn(2, 6, 4, generator=g, dtype=torch.float64, requires_grad=True)
    row = torch.randn(2, 1, 4, generator=g, dtype=torch.float64, requires_grad=True)
    base2, row2 = base.detach().clone().requires_grad_(True), row.detach().clone().requires_grad_(True)
    w = torch.randn(2, 6, 4, generator=g, dtype=torch.float64)
    ref = torch.cat([row2, (base2 * 2)[..., 1:, :]], dim=-2)
    out = _ReplaceFirstPosition.apply(base * 2, row)             # `base * 2` is a temporary, like a layer's output
    assert torch.equal(out, ref)
    (out * w).sum().backward()
    (ref * w).sum().backward()
    assert torch.equal(base.grad, base2.grad) and torch.equal(row.grad, row2.grad)


# ---------------------------------------------------------------- graphed decoding: static restatement of process_logits
@pytest.mark.parametrize('kwargs', [dict(), dict(temperature=0.0), dict(top_k=5), dict(top_p=1.0), dict(repetition_penalty=1.0),
                                    dict(top_k=7, top_p=0.5, temperature=0.7)])
def test_static_process_logits_equals_generation_state(kwargs):
    """core/decode.py restates GenerationState.process_logits with fixed shapes (device-side column counter, clamped
    penalty window).  On the CPU both run on the same logits with the same generator state: identical tokens."""
    import types

    from sparse_vae_b200.core import decode
    from sparse_vae_b200.core.generation import GenerationState

    torch.manual_seed(0)
    B, V, max_len = 5, 512, 700
    state = GenerationState(max_len, B, 1, 2, device=torch.device('cpu'), **kwargs)
    shadow = GenerationState(max_len, B, 1, 2, device=torch.device('cpu'), **kwargs)
    for step in range(1, 600):
        logits = torch.randn(B, V) * 3
        logits[:, 2] = -50.0                          # never the end token: the live batch stays complete
        if step in (3, 40, 599):                      # compare at a short, a medium and a > 512-token history
            dec = types.SimpleNamespace(state=shadow, ids=shadow.output_ids, column=torch.tensor([[shadow.current_index]]),
                                        window_offsets=torch.arange(-decode.PENALTY_WINDOW, 0)[None])
            rng = torch.get_rng_state()
            want_logits = logits.clone()
            state.process_logits(want_logits)
            want = state.output_ids[:, state.current_index - 1].clone()
            torch.set_rng_state(rng)
            got = decode.GraphedDecoder._process_logits(dec, logits.clone())
            assert torch.equal(got, want), (step, got, want)
            shadow.output_ids[:, shadow.current_index] = got
            shadow.current_index += 1
        else:                                         # fill the history with arbitrary (never end) tokens
            tok = torch.randint(3, V, (B,))
            for s in (state, shadow):
                s.output_ids[:, s.current_index] = tok
                s.current_index += 1


# ---------------------------------------------------------------- graphed decoding: ring KV cache vs the layout oracle
def _ring_slot(p, block, window):
    return p if p < block else block + (p - block) % (window * block)


def _slot_position(slot, pos, block, window):
    """Python statement of csrc/decode_attn.cu::slot_position."""
    if slot < block:
        return slot if slot <= pos else -1
    if pos < block:
        return -1
    ring = window * block
    q = block + (pos - block) - ((pos - block) - (slot - block)) % ring
    if q < block:
        return -1
    first_block = max(1, pos // block - (window - 1))
    return q if q >= first_block * block else -1


@pytest.mark.parametrize('window', [1, 2, 4, 7])
def test_ring_cache_shows_exactly_the_keys_of_the_layout_row(window):
    """Decoding position p must see the keys that row p's block of the causal include_cls layout allows (and, inside the
    diagonal block, only keys <= p): the ring cache of decode_attn_kernel holds exactly those, each in one slot."""
    from oracle import decoding as odec
    block, cache = 32, (window + 1) * 32
    length = 32 * 14
    stored = {}                                            # slot -> position last written there
    for p in range(length):
        stored[_ring_slot(p, block, window)] = p
        visible = {}
        for slot in range(cache):
            q = _slot_position(slot, p, block, window)
            if q >= 0:
                assert stored.get(slot) == q, (p, slot, q, stored.get(slot))      # the slot really holds that position
                visible[q] = slot
        want = odec.visible_positions(p, window, block)
        assert set(visible) == want, (p, sorted(set(visible) ^ want)[:8])
