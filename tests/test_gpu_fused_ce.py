"""Fused vocabulary head + cross-entropy (csrc/vocab_ce.cu, core/fused_ce.py) against the reference's
`robust_cross_entropy(Linear(hidden)[..., :-1, :], labels)` (core/language_model.py:161-170)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _reference_nll(hidden, weight, bias, labels):
    """Literal restatement of the reference: full logits, slice, chunked F.cross_entropy(ignore_index=0)."""
    logits = F.linear(hidden, weight, bias)[..., :-1, :]
    chunks = -(-logits.numel() // 2 ** 30)
    if chunks == 1:
        return F.cross_entropy(logits.flatten(end_dim=1), labels.flatten(), ignore_index=0)
    return torch.stack([F.cross_entropy(lo.flatten(end_dim=1), la.flatten(), ignore_index=0)
                        for lo, la in zip(logits.chunk(chunks, dim=-2), labels.chunk(chunks, dim=-1))]).mean()


def _setup(B, L, D, V, seed, pad_from=None):
    g = torch.Generator().manual_seed(seed)
    hidden = torch.randn(B, L, D, generator=g)
    weight = torch.randn(V, D, generator=g) * 0.05
    bias = torch.randn(V, generator=g) * 0.1
    labels = torch.randint(1, V, (B, L - 1), generator=g)
    if pad_from is not None:
        for b, p in enumerate(pad_from):
            labels[b, p:] = 0
    return hidden, weight, bias, labels


@pytest.mark.parametrize('V', [8192, 32768])
@pytest.mark.parametrize('row_chunk', [4096, 100])
def test_fused_ce_fp32_matches_float64_reference(V, row_chunk):
    from sparse_vae_b200.core.fused_ce import fused_vocab_nll
    B, L, D = 3, 130, 64
    hidden, weight, bias, labels = _setup(B, L, D, V, seed=V + row_chunk, pad_from=[129, 77, 5])
    hd, wd, bd = (t.double().requires_grad_(True) for t in (hidden, weight, bias))
    ref = _reference_nll(hd, wd, bd, labels)
    ref.backward()
    lin = torch.nn.Linear(D, V).cuda()
    with torch.no_grad():
        lin.weight.copy_(weight); lin.bias.copy_(bias)
    h = hidden.cuda().requires_grad_(True)
    loss = fused_vocab_nll(h, lin, labels.cuda(), row_chunk=row_chunk)
    (loss * 1.0).backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    for got, want in ((h.grad, hd.grad), (lin.weight.grad, wd.grad), (lin.bias.grad, bd.grad)):
        assert (got.double().cpu() - want).abs().max() <= 2e-4 * want.abs().max()     # TF32-free fp32 GEMMs
    # upstream gradient scaling
    h2 = hidden.cuda().requires_grad_(True)
    lin.zero_grad()
    (fused_vocab_nll(h2, lin, labels.cuda(), row_chunk=row_chunk) * 0.25).backward()
    torch.testing.assert_close(h2.grad, h.grad * 0.25, rtol=1e-6, atol=1e-9)


def test_fused_ce_multi_chunk_weighting_matches_reference():
    """More than 2**30 logits: the reference averages per-chunk means (chunks of unequal valid-token counts)."""
    from sparse_vae_b200.core.fused_ce import fused_vocab_nll
    B, L, D, V = 4, 8200, 32, 32768                     # 4 * 8199 * 32768 > 2**30 -> 2 chunks along the sequence
    hidden, weight, bias, labels = _setup(B, L, D, V, seed=11, pad_from=[8199, 6000, 4100, 300])
    lin = torch.nn.Linear(D, V).cuda()
    with torch.no_grad():
        lin.weight.copy_(weight); lin.bias.copy_(bias)
    h = hidden.cuda().requires_grad_(True)
    loss = fused_vocab_nll(h, lin, labels.cuda())
    loss.backward()
    h_ref = hidden.cuda().requires_grad_(True)
    w_ref, b_ref = (t.detach().clone().requires_grad_(True) for t in (lin.weight, lin.bias))
    ref = _reference_nll(h_ref, w_ref, b_ref, labels.cuda())          # plain ATen ops on the GPU, fp32
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert (h.grad - h_ref.grad).abs().max() <= 1e-3 * h_ref.grad.abs().max()
    assert (lin.weight.grad - w_ref.grad).abs().max() <= 1e-3 * w_ref.grad.abs().max()
    assert (lin.bias.grad - b_ref.grad).abs().max() <= 1e-3 * b_ref.grad.abs().max()


def test_fused_ce_autocast_matches_reference_sequence():
    from sparse_vae_b200.core.fused_ce import fused_vocab_nll
    B, L, D, V = 2, 513, 512, 32768
    hidden, weight, bias, labels = _setup(B, L, D, V, seed=5, pad_from=[512, 400])
    lin = torch.nn.Linear(D, V).cuda()
    with torch.no_grad():
        lin.weight.copy_(weight); lin.bias.copy_(bias)
    h = hidden.cuda().requires_grad_(True)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        loss = fused_vocab_nll(h, lin, labels.cuda())
    loss.backward()
    h_ref = hidden.cuda().requires_grad_(True)
    w_ref, b_ref = (t.detach().clone().requires_grad_(True) for t in (lin.weight, lin.bias))
    with torch.autocast('cuda', dtype=torch.bfloat16):
        ref = _reference_nll(h_ref, w_ref, b_ref, labels.cuda())
    ref.backward()
    assert loss.dtype == torch.float32
    # same rounded logits; this kernel keeps the log-softmax in fp32 (the reference's torch-1.9 autocast did too),
    # ATen 2.11 rounds the log-probabilities to bf16 -> agreement to ~1e-5..1e-4 of the mean, not bit for bit
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    assert h.grad.dtype == torch.float32
    assert (h.grad - h_ref.grad).abs().max() <= 1e-2 * h_ref.grad.abs().max()
    assert (lin.weight.grad - w_ref.grad).abs().max() <= 1e-2 * w_ref.grad.abs().max()
    assert (lin.bias.grad - b_ref.grad).abs().max() <= 1e-2 * b_ref.grad.abs().max()


def test_fused_ce_all_ignored_is_nan_like_reference():
    from sparse_vae_b200.core.fused_ce import fused_vocab_nll
    lin = torch.nn.Linear(32, 8192).cuda()
    h = torch.randn(1, 9, 32, device='cuda')
    labels = torch.zeros(1, 8, dtype=torch.long, device='cuda')
    assert torch.isnan(fused_vocab_nll(h, lin, labels))
    assert torch.isnan(_reference_nll(h, lin.weight, lin.bias, labels))


def test_fused_ce_rejects_cpu():
    from sparse_vae_b200.core.fused_ce import fused_vocab_nll
    with pytest.raises(ValueError):
        fused_vocab_nll(torch.randn(1, 4, 8), torch.nn.Linear(8, 8192), torch.ones(1, 3, dtype=torch.long))


def test_fused_ce_fp16_autocast():
    """fp16 (the reference's own `precision=16`): same contract as bf16, tighter mantissa."""
    from sparse_vae_b200.core.fused_ce import fused_vocab_nll
    B, L, D, V = 2, 257, 256, 16384
    hidden, weight, bias, labels = _setup(B, L, D, V, seed=9, pad_from=[256, 100])
    lin = torch.nn.Linear(D, V).cuda()
    with torch.no_grad():
        lin.weight.copy_(weight); lin.bias.copy_(bias)
    h = hidden.cuda().requires_grad_(True)
    with torch.autocast('cuda', dtype=torch.float16):
        loss = fused_vocab_nll(h, lin, labels.cuda())
    loss.backward()
    hd, wd, bd = (t.double().requires_grad_(True) for t in (hidden, weight, bias))
    ref = _reference_nll(hd, wd, bd, labels)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-3 * abs(ref.item())
    assert (h.grad.double().cpu() - hd.grad).abs().max() <= 1e-2 * hd.grad.abs().max()
    assert (lin.weight.grad.double().cpu() - wd.grad).abs().max() <= 1e-2 * wd.grad.abs().max()
    assert (lin.bias.grad.double().cpu() - bd.grad).abs().max() <= 1e-2 * bd.grad.abs().max()


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
@pytest.mark.parametrize('rows,V', [(1, 8192), (37, 16384), (149, 32768), (1000, 32768), (300, 24576)])
def test_streamed_kernel_equals_register_resident_kernel_and_float64(dtype, rows, V):
    """The streamed 16-bit kernel (persistent CTAs, shared-memory ring, bulk copies) against the one-row-per-CTA
    kernel (same arithmetic, another summation tree) and against float64."""
    from sparse_vae_b200 import _native as N
    g = torch.Generator().manual_seed(rows + V)
    logits = (torch.randn(rows, V + 8, generator=g) * 3).to('cuda', dtype)[:, :V]          # row stride V + 8
    labels = torch.randint(1, V, (rows,), generator=g).cuda()
    weight = torch.rand(rows, generator=g).cuda()
    weight[::7] = 0.0                                                                         # ignored rows
    outs = []
    for variant in (1, 3):                     # write_grad bit 1: the register-resident kernel
        buf = torch.empty(rows, V + 8, device='cuda', dtype=dtype)
        buf[:, :V] = logits
        buf[:, V:] = 7.0                                                                      # must stay untouched
        nll = torch.full((rows,), -1.0, device='cuda')
        N.check(N.lib.svae_vocab_ce(buf.data_ptr(), N.svae_dtype(dtype), rows, V, buf.stride(0), labels.data_ptr(),
                                    weight.data_ptr(), nll.data_ptr(), variant, N.current_stream(buf.device)), 'svae_vocab_ce')
        torch.cuda.synchronize()
        assert (buf[:, V:] == 7.0).all()
        outs.append((buf[:, :V].clone(), nll))
    # (different summation trees for the row sum: a last-bit difference of the sum moves a rounding now and then)
    assert (outs[0][0] != outs[1][0]).float().mean().item() < 1e-3
    assert (outs[0][1] - outs[1][1]).abs().max().item() <= 1e-5
    x = logits.double()
    lse = torch.logsumexp(x, -1)
    want_nll = torch.where(weight == 0, torch.zeros_like(lse), lse - x.gather(1, labels[:, None])[:, 0])
    assert (outs[0][1].double() - want_nll).abs().max() <= 1e-5 * want_nll.abs().max() + 1e-5
    grad = weight.double()[:, None] * (torch.softmax(x, -1) - torch.nn.functional.one_hot(labels, V))
    tol = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    assert ((outs[0][0].double() - grad).abs() <= tol * grad.abs() + 1e-7).all()
    # forward only: the logits stay as they are
    buf = logits.clone().contiguous()
    nll = torch.empty(rows, device='cuda')
    N.check(N.lib.svae_vocab_ce(buf.data_ptr(), N.svae_dtype(dtype), rows, V, buf.stride(0), labels.data_ptr(),
                                weight.data_ptr(), nll.data_ptr(), 0, N.current_stream(buf.device)), 'svae_vocab_ce')
    assert torch.equal(buf, logits.contiguous()) and torch.equal(nll, outs[0][1])
