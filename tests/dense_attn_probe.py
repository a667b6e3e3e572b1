"""Probe: the encoder's dense attention (64 learned queries over 4096 keys) -- explicit softmax vs F.scaled_dot_product_attention."""
import torch
import torch.nn.functional as F

dev = torch.device('cuda')
B, H, Lq, Lk, Dh = 16, 8, 64, 4096, 64
torch.manual_seed(0)
qf = torch.randn(B, Lq, H * Dh, device=dev, dtype=torch.bfloat16, requires_grad=True)
kf = torch.randn(B, Lk, H * Dh, device=dev, dtype=torch.bfloat16, requires_grad=True)
vf = torch.randn(B, Lk, H * Dh, device=dev, dtype=torch.bfloat16, requires_grad=True)
pad = torch.zeros(B, Lk, dtype=torch.bool, device=dev)
pad[3, 3000:] = True


def split(t):
    return t.unflatten(-1, (H, -1)).transpose(-2, -3)


def manual(mask):
    q, k, v = split(qf), split(kf), split(vf)
    scores = q @ k.transpose(-1, -2) * Dh ** -0.5
    if mask is not None:
        scores = scores - mask[:, None, None, :] * 1e7
    return (scores.softmax(dim=-1) @ v).transpose(-2, -3).flatten(-2)


def sdpa(mask):
    q, k, v = split(qf), split(kf), split(vf)
    bias = None if mask is None else (mask[:, None, None, :] * -1e7).to(q.dtype)
    return F.scaled_dot_product_attention(q, k, v, attn_mask=bias).transpose(-2, -3).flatten(-2)


def timeit(fn, mask):
    for _ in range(3):
        out = fn(mask)
        out.sum().backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = fn(mask)
        out.backward(torch.ones_like(out))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10


ctx = torch.autocast('cuda', dtype=torch.bfloat16)
ctx.__enter__()
for mask in (None, pad):
    a = manual(mask).float()
    b = sdpa(mask).float()
    print('mask' if mask is not None else 'no mask', 'max rel diff', ((a - b).abs().max() / a.abs().max()).item(),
          'manual ms', round(timeit(manual, mask), 3), 'sdpa ms', round(timeit(sdpa, mask), 3))
