"""Token-embedding weight gradient (csrc/embedding.cu, core/embedding.py) against nn.Embedding."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('n,vocab,d,skew', [(7, 16, 4, False), (4096, 1000, 64, True), (65536, 32768, 512, False), (5000, 32768, 512, True)])
def test_embedding_backward_matches_float64_and_is_deterministic(dtype, n, vocab, d, skew):
    from sparse_vae_b200 import _native as N
    g = torch.Generator().manual_seed(n + d)
    ids = torch.randint(0, vocab, (n,), generator=g)
    if skew:
        ids[torch.rand(n, generator=g) < 0.6] = 3          # one token takes most positions, many tokens never occur
    grad = torch.randn(n, d, generator=g).to('cuda', dtype)
    ids = ids.cuda()
    sorted_ids, perm = torch.sort(ids, stable=True)
    bounds = torch.searchsorted(sorted_ids, torch.arange(vocab + 1, device='cuda'))
    outs = []
    for _ in range(2):
        dw = torch.full((vocab, d), 7.0, device='cuda')
        N.check(N.lib.svae_embedding_bwd(grad.data_ptr(), N.svae_dtype(dtype), bounds.data_ptr(), perm.data_ptr(), n, vocab, d,
                                         dw.data_ptr(), N.current_stream(dw.device)), 'svae_embedding_bwd')
        outs.append(dw)
    assert torch.equal(outs[0], outs[1])
    ref = torch.zeros(vocab, d, dtype=torch.float64, device='cuda').index_add_(0, ids, grad.double())
    assert (outs[0].double() - ref).abs().max() <= 1e-5 * max(ref.abs().max().item(), 1.0)


def test_embedding_module_matches_nn_embedding():
    from sparse_vae_b200.core.embedding import Embedding
    torch.manual_seed(4)
    ours, ref = Embedding(32768, 512).cuda(), torch.nn.Embedding(32768, 512).cuda()
    ref.load_state_dict(ours.state_dict())
    assert set(ours.state_dict()) == {'weight'}
    ids = torch.randint(0, 32768, (16, 1024), device='cuda')
    dy = torch.randn(16, 1024, 512, device='cuda')
    ya, yb = ours(ids), ref(ids)
    assert torch.equal(ya, yb) and type(ya.grad_fn).__name__ == '_EmbeddingFnBackward'
    ya.backward(dy)
    yb.backward(dy)
    assert (ours.weight.grad - ref.weight.grad).abs().max() <= 1e-5 * ref.weight.grad.abs().max()
    with torch.no_grad():
        assert torch.equal(ours(ids), yb)                    # no-grad calls take the library path
