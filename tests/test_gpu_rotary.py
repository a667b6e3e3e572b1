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
