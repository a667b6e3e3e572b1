"""GPU: token-by-token decoding -- `svae_decode_attn` against the CPU oracle's block-sparse
attention, and the CUDA-graphed sampler against the teacher-forced block-sparse decoder on the tokens it produced.

Tolerance: 1e-2 relative for 16-bit, 1e-4 for fp32 (BASELINE.json north_star), relative = max|a-b| / max|b|."""
import sys
from pathlib import Path

import pytest
import torch

sys.path.insert(0, str(Path(__file__).parent))
sys.path.insert(0, str(Path(__file__).parent / 'golden'))
import make_golden as mg  # noqa: E402
from oracle import decoding as odec  # noqa: E402
from util import oracle_attention, rel_err  # noqa: E402

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.bfloat16: 1e-2, torch.float16: 1e-2}


def _decode_all(q, k, v, cos, sin, H, window, start=0):
    """Feeds rows start..T-1 of q/k/v [B, T, D] one at a time; the caches persist between launches like in sample()."""
    from sparse_vae_b200 import _native as N
    B, T, D = q.shape
    dev = q.device
    C = (window + 1) * 32
    key_cache = torch.zeros(B, C, D, dtype=q.dtype, device=dev)
    value_cache = torch.zeros(B, C, D, dtype=q.dtype, device=dev)
    position = torch.zeros(1, dtype=torch.int32, device=dev)
    outs = torch.empty(B, T, D, dtype=q.dtype, device=dev)
    qkv = torch.empty(B, 3 * D, dtype=q.dtype, device=dev)
    for t in range(start, T):
        qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:] = q[:, t], k[:, t], v[:, t]
        out = torch.empty(B, D, dtype=q.dtype, device=dev)
        position.fill_(t)
        N.check(N.lib.svae_decode_attn(qkv[:, :D].data_ptr(), qkv[:, D:2 * D].data_ptr(), qkv[:, 2 * D:].data_ptr(),
                                       cos.data_ptr(), sin.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(),
                                       out.data_ptr(), position.data_ptr(), B, H, D // H, window, 32, cos.shape[0], 3 * D,
                                       N.svae_dtype(q.dtype), float((D // H) ** -0.5), N.current_stream(dev)),
                'svae_decode_attn')
        outs[:, t] = out
    return outs


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize('H,Dh,window,T', [(8, 64, 4, 416), (8, 32, 4, 224), (2, 64, 2, 160), (2, 64, 7, 352), (4, 32, 1, 96)])
def test_decode_steps_match_oracle_block_sparse_rows(dtype, H, Dh, window, T):
    """Row t of the causal block-sparse attention over the whole sequence == decoding step t over the ring cache
    (covers the first block, the fill phase, the first wrap-around and several evictions)."""
    import sparse_vae_b200 as sv
    from sparse_vae_b200.core.attention import _rotary_tables, encode_position_rotary
    dev = torch.device('cuda')
    B, D = 3, H * Dh
    g = torch.Generator(device='cpu').manual_seed(1234 + T)
    q, k, v = (torch.randn(B, T, D, generator=g).to(dev, dtype) for _ in range(3))
    max_pos = 2 * window * 32
    with torch.autocast('cuda', dtype=dtype, enabled=dtype != torch.float32):
        cos, sin = _rotary_tables(T, D // 2, 0, max_pos, dtype, dev)
        q_rot, k_rot = encode_position_rotary(q, 0, max_pos), encode_position_rotary(k, 0, max_pos)
    cos, sin = cos.float().contiguous(), sin.float().contiguous()
    out = _decode_all(q, k, v, cos, sin, H, window)
    cfg = sv.SparseAttention(window_size=window, num_heads=H)
    split = lambda t: t.to(dtype).unflatten(-1, (H, Dh)).transpose(1, 2)      # noqa: E731  [B, H, T, Dh]
    ref = oracle_attention(split(q_rot), split(k_rot), split(v), cfg)
    err = rel_err(split(out), ref)
    assert err <= TOL[dtype], err
    # and against the training kernel itself (same rounded inputs)
    if dtype != torch.float32:
        fused = cfg(split(q_rot), split(k_rot), split(v))
        assert rel_err(split(out), fused.double().cpu()) <= TOL[dtype]


def test_decode_rejects_unsupported_shapes():
    from sparse_vae_b200 import _native as N
    assert N.lib.svae_decode_attn_supported(64, 4, 32) == 1
    assert N.lib.svae_decode_attn_supported(48, 4, 32) == 0 and N.lib.svae_decode_attn_supported(64, 15, 32) == 0
    t = torch.zeros(64, device='cuda')
    rc = N.lib.svae_decode_attn(t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(),
                                t.data_ptr(), t.data_ptr(), t.data_ptr(), 1, 1, 48, 4, 32, 8, 48, 0, 1.0, None)
    assert rc == -2 and b'head_dim' in N.lib.svae_last_error()


def _small_model(dev, scale=1.0, **over):
    import sparse_vae_b200 as sv
    from sparse_vae_b200.core.lightning_shim import to_attrdict
    case = mg.MODEL_CASE
    hp = dict(d_model=case['d_model'], num_layers=case['num_layers'], num_heads=case['num_heads'],
              attn_window_size=case['window'], latent_depth=case['latent'])
    hp.update(over)
    torch.manual_seed(11)
    model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams(**hp))).to(dev).eval()
    model.initialize_weights()
    with torch.no_grad():                   # scale > 1: sharper logits than the 0.02 init gives (decoding is no coin flip)
        for name, p in model.named_parameters():
            if 'layer_norm' not in name and scale != 1.0:
                p.mul_(scale)
    model.hparams.kl_weight = 1.0
    model.start_token, model.end_token = 1, 2
    return sv, model


@pytest.mark.parametrize('length', [80, 300])
def test_graphed_sampler_logits_match_teacher_forced_decoder(length):
    """Every step of the graphed sampler must produce the logits the block-sparse training path gives at that position
    for the same tokens and z: ties the decode kernel, the ring cache and the device-side counters to the sparse
    kernel.  (The sampled ids themselves depend on the RNG stream and are checked for range / shape only.)"""
    from sparse_vae_b200.core import decode
    dev = torch.device('cuda')
    sv, model = _small_model(dev)
    B = 3
    z = torch.randn(B, 1, model.hparams.latent_depth, device=dev)
    decode.TRACE = []
    try:
        with torch.no_grad():
            ids = model.sample(length, B, z=z)
        trace = decode.TRACE
    finally:
        decode.TRACE = None
    assert ids.shape == (B, length - 1) and ((ids >= 0) & (ids < 2 ** 15)).all()
    assert len(trace) >= 1, "graphed decoder did not run"
    # teacher-forced pass over [start] + generated tokens, padded to a multiple of the block size
    full = torch.cat([torch.full((B, 1), model.start_token, device=dev), ids], dim=1)
    L = (full.shape[1] + 31) // 32 * 32
    tokens = torch.zeros(B, L, dtype=torch.long, device=dev)
    tokens[:, :full.shape[1]] = full
    with torch.no_grad(), torch.autocast('cuda'):
        ref = model.reconstruct(model.input_layer(tokens), z).float()        # [B, L, V]; causal: padding cannot leak back
    worst = 0.0
    for step, (rows, logits) in enumerate(trace):
        t = step + 1                        # graphed step `step` is fed the token in column t and predicts column t+1
        want = ref[rows, t]
        worst = max(worst, ((logits.float() - want).abs().max() / want.abs().max()).item())
    assert worst <= 2e-2, worst             # fp16 autocast on both sides, different GEMM shapes


def test_graphed_sampler_drops_finished_samples_like_the_reference():
    """A sample that draws the end token leaves the batch: its row stays zero afterwards, the others keep going."""
    from sparse_vae_b200.core import decode
    dev = torch.device('cuda')
    sv, model = _small_model(dev)
    with torch.no_grad():
        model.output_layer[3].bias[model.end_token] += 6.0                  # end token likely, not certain
        torch.manual_seed(3)
        decode.TRACE = []
        try:
            ids = model.sample(200, 16)
            trace = decode.TRACE
        finally:
            decode.TRACE = None
    live_counts = [rows.numel() for rows, _ in trace]
    assert live_counts == sorted(live_counts, reverse=True) and live_counts[-1] < 16, live_counts
    for row in ids.tolist():
        if model.end_token in row:
            end = row.index(model.end_token)
            assert all(tok == 0 for tok in row[end + 1:])


def test_greedy_graphed_sampler_equals_module_sampler():
    """temperature 0: no RNG involved, so the graphed sampler and the module-by-module (reference-order) sampler must
    pick the same tokens wherever the module sampler's top-2 logit gap exceeds the 16-bit noise."""
    from sparse_vae_b200 import _native as N
    dev = torch.device('cuda')
    sv, model = _small_model(dev, scale=4.0)
    z = torch.randn(4, 1, model.hparams.latent_depth, device=dev)
    with torch.no_grad():
        fast = model.sample(72, 4, z=z, temperature=0.0)
        N.FUSED_EXTRAS = False
        try:
            slow = model.sample(72, 4, z=z, temperature=0.0)
        finally:
            N.FUSED_EXTRAS = True
    same_prefix = (fast == slow).long().cumprod(dim=1).sum(dim=1)
    assert (same_prefix >= 8).all(), same_prefix              # decoding agrees until a near-tie (if any) flips a token
    assert (same_prefix == fast.shape[1]).float().mean() >= 0.5, same_prefix


# ---------------------------------------------------------------- one-launch nucleus sampler (csrc/sampling.cu)
def _run_sampler(logits, uniforms, ids=None, column=1, alive=None, penalty=1.0, temperature=1.0, top_p=0.9, end_token=2):
    from sparse_vae_b200 import _native as N
    B, V = logits.shape
    dev = logits.device
    ids = torch.zeros(B, 8, dtype=torch.long, device=dev) if ids is None else ids
    alive = torch.ones(B, dtype=torch.bool, device=dev) if alive is None else alive
    finished = torch.zeros(1, dtype=torch.int32, device=dev)
    col = torch.full((1, 1), column, dtype=torch.long, device=dev)
    N.check(N.lib.svae_sample_top_p(logits.data_ptr(), N.svae_dtype(logits.dtype), B, V, ids.data_ptr(), ids.stride(0),
                                    col.data_ptr(), uniforms.data_ptr(), alive.data_ptr(), finished.data_ptr(), 512,
                                    penalty, temperature, top_p, end_token, N.current_stream(dev)), 'svae_sample_top_p')
    return ids, alive, int(finished.item())


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
@pytest.mark.parametrize('V,top_p,spread', [(4096, 0.9, 3.0), (8192, 0.5, 1.0), (1000, 0.97, 6.0), (4096, 1.0, 2.0)])
def test_sampler_draws_the_inverse_cdf_token_of_the_reference_nucleus(dtype, V, top_p, spread):
    """V <= 8192: the kernel's sampling order is the index order, so the token for a given uniform is known exactly."""
    dev = torch.device('cuda')
    g = torch.Generator(device='cpu').manual_seed(V + int(100 * top_p))
    row = (torch.randn(V, generator=g) * spread).to(dtype)
    B = 1024
    logits = row.to(dev)[None].repeat(B, 1).contiguous()
    uniforms = torch.rand(B, generator=g).to(dev)
    ids, _, _ = _run_sampler(logits, uniforms, top_p=top_p, end_token=-1)
    got = ids[:, 1].cpu()
    w = odec.nucleus_weights(row.double(), top_p)
    # the reference's own rule (sort, softmax, cumsum > top_p masked, first kept) keeps the same set up to boundary ties
    ref_keep = odec.nucleus_keep_reference(row, top_p)
    boundary = row.float()[w > 0].min()
    differs = (ref_keep != (w > 0))
    assert (row.float()[differs] == boundary).all() and differs.sum() <= 2 + (row.float() == boundary).sum()
    want = odec.inverse_cdf_token(w, uniforms.cpu())
    mismatch = (got != want)
    assert mismatch.float().mean() <= 0.005, (got[mismatch][:8], want[mismatch][:8])
    assert (w[got] > 0).all()                      # never a token outside the nucleus


def test_sampler_full_vocab_frequencies():
    """V = 32768 (four register chunks per thread): empirical frequencies of the likeliest tokens vs the nucleus."""
    dev = torch.device('cuda')
    g = torch.Generator(device='cpu').manual_seed(9)
    V, B = 32768, 8192
    row = (torch.randn(V, generator=g) * 2.5).to(torch.float16)
    logits = row.to(dev)[None].repeat(B, 1).contiguous()
    uniforms = torch.rand(B, generator=g).to(dev)
    ids, _, _ = _run_sampler(logits, uniforms, top_p=0.9, end_token=-1)
    got = ids[:, 1].cpu()
    w = odec.nucleus_weights(row.double(), 0.9)
    p = w / w.sum()
    assert (w[got] > 0).all()
    counts = torch.bincount(got, minlength=V).double()
    top = p.topk(12).indices
    sigma = (B * p[top] * (1 - p[top])).sqrt()
    assert ((counts[top] - B * p[top]).abs() <= 5 * sigma + 1).all(), (counts[top], B * p[top])
    # inverse CDF from evenly spread uniforms covers the nucleus evenly: total variation stays small
    assert 0.5 * (counts / B - p).abs().sum() <= 0.5


def test_sampler_penalty_temperature_and_bookkeeping():
    dev = torch.device('cuda')
    V = 32768
    logits = torch.full((4, V), -8.0, dtype=torch.float16, device=dev)
    logits[:, 100], logits[:, 200], logits[:, 300] = 5.0, 4.5, -7.0
    ids = torch.zeros(4, 8, dtype=torch.long, device=dev)
    ids[0, :3] = torch.tensor([1, 100, 7])            # row 0 has generated token 100 -> 5.0 / 1.2 < 4.5
    ids[1, :3] = torch.tensor([1, 300, 7])            # negative logits are multiplied: -7 * 1.2, irrelevant for the argmax
    ids[2, :3] = torch.tensor([1, 100, 7])            # dead row: nothing written
    ids[3, :3] = torch.tensor([1, 9, 200])            # row 3 draws 100 = its end token
    alive = torch.tensor([True, True, False, True], device=dev)
    uniforms = torch.full((4,), 0.3, device=dev)
    out, alive, finished = _run_sampler(logits, uniforms, ids=ids, column=3, alive=alive, penalty=1.2, top_p=0.05,
                                        end_token=100)
    # row 3: token 200 was generated -> 4.5 / 1.2 = 3.75, so 100 wins and ends the sample
    assert out[:, 3].tolist() == [200, 100, 0, 100]
    assert alive.tolist() == [True, False, False, False] and finished == 2
    # temperature: dividing by 0.01 makes the distribution a point mass even with top_p = 1
    out, _, _ = _run_sampler(logits, torch.rand(4, device=dev), column=1, temperature=0.01, top_p=1.0, end_token=-1)
    assert out[:, 1].tolist() == [100] * 4


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize('rows,n', [(1, 128), (256, 512), (37, 1024)])
def test_residual_layernorm_equals_add_then_layernorm(dtype, rows, n):
    """x += h; y = LN(x): bit-identical to the separate fp32 add and the LayerNorm kernel it replaces while decoding."""
    from sparse_vae_b200 import _native as N
    from sparse_vae_b200.core.layer_norm import LayerNorm
    dev = torch.device('cuda')
    torch.manual_seed(rows + n)
    norm = LayerNorm(n).to(dev)
    with torch.no_grad():
        norm.weight.normal_(1, 0.2)
        norm.bias.normal_(0, 0.2)
    x = torch.randn(rows, 1, n, device=dev) * 2
    h = torch.randn(rows, 1, n, device=dev).to(dtype)
    want_x = x + h
    with torch.no_grad(), torch.autocast('cuda', dtype=dtype, enabled=dtype != torch.float32):
        want_y = norm(want_x)
    y = torch.empty_like(h)
    N.check(N.lib.svae_residual_layernorm(x.data_ptr(), h.data_ptr(), N.svae_dtype(dtype), norm.weight.data_ptr(),
                                          norm.bias.data_ptr(), rows, n, norm.eps, y.data_ptr(), N.svae_dtype(dtype),
                                          x.data_ptr(), None, None, N.current_stream(dev)), 'svae_residual_layernorm')
    assert torch.equal(x, want_x) and want_y.dtype == dtype and torch.equal(y, want_y)
