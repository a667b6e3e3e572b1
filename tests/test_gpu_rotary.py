"""Single-launch rotary kernel (csrc/rotary.cu) against the reference's op sequence (core/attention.py:194-208),
bit for bit in every dtype, forward and backward."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(x, start, max_pos):
    """Literal restatement of the reference function with plain torch ops."""
    half = x.shape[-1] // 2
    freq = torch.arange(half, dtype=x.dtype, device=x.device)
    pos = torch.arange(start, start + x.shape[-2], dtype=x.dtype, device=x.device)
    angle = pos[:, None] * (max_pos ** (-freq / half))
    cos, sin = angle.cos(), angle.sin()
    pairs = x.unflatten(-1, (half, 2))
    even, odd = pairs[..., 0], pairs[..., 1]
    return torch.stack((even * cos - odd * sin, odd * cos + even * sin), dim=-1).flatten(-2)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize('shape,start,max_pos', [((2, 512, 256), 0, 256), ((3, 1000, 512), 0, 256), ((4, 1, 512), 37, 256),
                                                 ((2, 4096, 512), 0, 10000), ((1, 64, 8), 0, 64)])
def test_rotary_bit_exact(dtype, shape, start, max_pos):
    from sparse_vae_b200.core.attention import encode_position_rotary
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g).to('cuda', dtype)
    dy = torch.randn(*shape, generator=g).to('cuda', dtype)
    a = x.clone().requires_grad_(True)
    b = x.clone().requires_grad_(True)
    ya = encode_position_rotary(a, start, max_pos)
    yb = _reference(b, start, max_pos)
    assert torch.equal(ya, yb)
    ya.backward(dy)
    yb.backward(dy)
    assert torch.equal(a.grad, b.grad)


def test_rotary_matches_cpu_path():
    from sparse_vae_b200.core.attention import encode_position_rotary
    x = torch.randn(2, 96, 64)
    y_cpu = encode_position_rotary(x, 0, 256)
    y_gpu = encode_position_rotary(x.cuda(), 0, 256).cpu()
    assert (y_cpu - y_gpu).abs().max() <= 2e-6           # cos/sin of the two devices differ in the last bits


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
@pytest.mark.parametrize('shape,max_pos', [((2, 4096, 512), 256), ((3, 600, 256), 10000)])
def test_rotary_under_autocast_bit_exact(dtype, shape, max_pos):
    """Under autocast the reference promotes to fp32 (fp32 cos/sin from the autocast `pow`) and the consumer casts
    the result to the autocast dtype; the kernel must produce exactly that tensor and exactly autograd's gradient."""
    from sparse_vae_b200.core.attention import encode_position_rotary
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g).to('cuda', dtype)
    dy = torch.randn(*shape, generator=g).to('cuda', dtype)
    a = x.clone().requires_grad_(True)
    b = x.clone().requires_grad_(True)
    with torch.autocast('cuda', dtype=dtype):
        ya = encode_position_rotary(a, 0, max_pos)
        ref32 = _reference(b, 0, max_pos)
        assert ref32.dtype == torch.float32 and ya.dtype == dtype
        yb = ref32.to(dtype)                                   # what the consuming matmul / sparse kernel sees
    assert torch.equal(ya, yb)
    ya.backward(dy)
    yb.backward(dy)
    assert torch.equal(a.grad, b.grad)


@pytest.mark.parametrize('dtype,table', [(torch.bfloat16, torch.float32), (torch.float16, torch.float32),
                                         (torch.bfloat16, torch.bfloat16), (torch.float32, torch.float32)])
@pytest.mark.parametrize('B,L,d', [(1, 32, 8), (2, 96, 520), (3, 4096, 512)])
def test_rotary_pair_equals_two_launches_and_colsum(dtype, table, B, L, d):
    """`svae_rotary_pair`: both outputs bit-identical to `svae_rotary`, the column sums bit-identical to `svae_colsum`."""
    from sparse_vae_b200 import _native as N
    from sparse_vae_b200.core.linear import colsum, rotary_pair
    g = torch.Generator().manual_seed(B * L + d)
    a, b = (torch.randn(B, L, d, generator=g).to('cuda', dtype) for _ in range(2))
    ang = torch.rand(L, d // 2, generator=g) * 6.0
    cos, sin = ang.cos().to('cuda', table).contiguous(), ang.sin().to('cuda', table).contiguous()

    def single(x, conj):
        out = torch.empty_like(x)
        N.check(N.lib.svae_rotary(x.data_ptr(), cos.data_ptr(), sin.data_ptr(), out.data_ptr(), N.svae_dtype(dtype),
                                  N.svae_dtype(table), B * L, L, d, conj, N.current_stream(x.device)), 'svae_rotary')
        return out

    for conj in (0, 1):
        oa, ob, sa, sb = rotary_pair(a, b, cos, sin, conj=bool(conj), want_colsum=True)
        assert torch.equal(oa, single(a, conj)) and torch.equal(ob, single(b, conj))
        assert torch.equal(sa, colsum(oa.view(-1, d))) and torch.equal(sb, colsum(ob.view(-1, d)))
        pa, pb, none_a, none_b = rotary_pair(a, b, cos, sin, conj=bool(conj))
        assert torch.equal(pa, oa) and torch.equal(pb, ob) and none_a is None and none_b is None


def test_qkv_rotary_node_matches_separate_ops():
    """core/linear.py `qkv_rotary` against linear3 + two encode_position_rotary calls: same forward bits, same gradients."""
    from sparse_vae_b200.core.attention import Attention
    torch.manual_seed(11)
    att = Attention(512, 8, causal=True, sparse=4).cuda()
    x = torch.randn(2, 1024, 512, device='cuda')
    dy = torch.randn(2, 1024, 512, device='cuda').to(torch.bfloat16)
    res = {}
    from sparse_vae_b200.core import attention as A
    for fused in (True, False):
        xin = x.clone().requires_grad_(True)
        att.zero_grad()
        orig = A.qkv_rotary
        if not fused:
            A.qkv_rotary = lambda *a, **k: None
        try:
            with torch.autocast('cuda', dtype=torch.bfloat16):
                y = att(xin, xin, xin)
            y.backward(dy)
        finally:
            A.qkv_rotary = orig
        res[fused] = (y.detach().clone(), xin.grad.clone(), {n: p.grad.clone() for n, p in att.named_parameters() if p.grad is not None})
    assert torch.equal(res[True][0], res[False][0])
    assert torch.equal(res[True][1], res[False][1])
    assert res[True][2].keys() == res[False][2].keys()
    for n in res[True][2]:
        assert torch.equal(res[True][2][n], res[False][2][n]), n


def test_qkv_joint_gemm_path_matches_separate_ops():
    """With the 16-bit weight shadows active the three projections run as ONE GEMM forward and one input- / one
    weight-gradient GEMM backward on the attention backward's joint [dq | dk | dv] buffer: same numbers as the separate
    ops up to the GEMMs' summation order."""
    from sparse_vae_b200.core import attention as A
    from sparse_vae_b200.core.attention import Attention
    from sparse_vae_b200.core.linear import WeightShadows
    torch.manual_seed(12)
    att = Attention(512, 8, causal=True, sparse=4).cuda()
    shadows = WeightShadows(att)
    x = torch.randn(2, 2048, 512, device='cuda')
    dy = torch.randn(2, 2048, 512, device='cuda').to(torch.bfloat16)
    res = {}
    for fused in (True, False):
        xin = x.clone().requires_grad_(True)
        att.zero_grad()
        orig = A.qkv_rotary
        if not fused:
            A.qkv_rotary = lambda *a, **k: None
        try:
            with torch.autocast('cuda', dtype=torch.bfloat16), shadows.step():
                y = att(xin, xin, xin)
            y.backward(dy)
        finally:
            A.qkv_rotary = orig
        grads = {n: p.grad.clone() for n, p in att.named_parameters() if p.grad is not None}
        if fused:      # the joint path ran: the three weight gradients are row blocks of one [3 d, d] GEMM result
            gq, gk, gv = (getattr(att, n).weight.grad for n in ('q_linear', 'k_linear', 'v_linear'))
            assert gk.data_ptr() == gq.data_ptr() + gq.numel() * 4 and gv.data_ptr() == gk.data_ptr() + gk.numel() * 4
        res[fused] = (y.detach().float(), xin.grad.clone(), grads)

    def close(a, b, tol):
        return (a.float() - b.float()).abs().max().item() <= tol * b.float().abs().max().item()

    assert close(res[True][0], res[False][0], 1e-2)
    assert close(res[True][1], res[False][1], 1e-2)
    assert res[True][2].keys() == res[False][2].keys()
    for n in res[True][2]:
        assert close(res[True][2][n], res[False][2][n], 1e-2), n
