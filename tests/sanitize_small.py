"""Small invocation of every library kernel (for compute-sanitizer --tool memcheck; not a test, not a bench)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import sparse_vae_b200 as sv  # noqa: E402
from sparse_vae_b200.core.attention import encode_position_rotary  # noqa: E402
from sparse_vae_b200.core.conditional_gaussian import fused_bottleneck  # noqa: E402
from sparse_vae_b200.core.fused_ce import fused_vocab_nll  # noqa: E402
from sparse_vae_b200.core.layer_norm import LayerNorm  # noqa: E402
from sparse_vae_b200.core.linear import Linear  # noqa: E402
from sparse_vae_b200.core.rectified_adam import RAdam  # noqa: E402
from sparse_vae_b200.fused_optim import FusedGradClipper  # noqa: E402
from util import make_padding, make_qkv  # noqa: E402

dev = torch.device('cuda')
import os
for persistent in ('1', '0'):
    os.environ['SVAE_ATTN_PERSISTENT'] = persistent
    for (B, L, w, lengths) in ((2, 160, 4, [160, 123]), (1, 640, 8, [600]), (3, 1024, 4, None), (40, 512, 4, None)):
        cfg = sv.SparseAttention(window_size=w)
        q, k, v = make_qkv(B, 8, L, 64, torch.bfloat16, dev, seed=L, requires_grad=True)
        pad = make_padding(B, L, lengths, dev) if lengths else None
        out = cfg(q, k, v, key_padding_mask=pad * -1e7 if pad is not None else None)
        out.backward(torch.randn_like(out))
q, k, v = make_qkv(1, 8, 128, 32, torch.float32, dev, seed=1, requires_grad=True)
sv.SparseAttention()(q, k, v).backward(torch.ones(1, 8, 128, 32, device=dev))          # exact kernels
fused_bottleneck((torch.randn(16, 1, 128, device=dev) * 0.5).to(torch.bfloat16), torch.full((16,), 4096, device=dev))
x = torch.randn(2, 512, 512, device=dev, dtype=torch.bfloat16, requires_grad=True)
with torch.autocast('cuda', dtype=torch.bfloat16):
    encode_position_rotary(x, 0, 256).float().sum().backward()
ln, lin = LayerNorm(512).to(dev), Linear(512, 512).to(dev)
xf = torch.randn(2048, 512, device=dev, requires_grad=True)
with torch.autocast('cuda', dtype=torch.bfloat16):
    lin(ln(xf)).float().sum().backward()
head = torch.nn.Linear(64, 32768).to(dev)
hid = torch.randn(2, 70, 64, device=dev, requires_grad=True)
lab = torch.randint(1, 32768, (2, 69), device=dev)
with torch.autocast('cuda', dtype=torch.bfloat16):
    fused_vocab_nll(hid, head, lab).backward()
fused_vocab_nll(hid.detach(), head, lab)
params = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in ((512, 512), (7,), (65536 + 3,))]
for p_ in params:
    p_.grad = torch.randn_like(p_)
FusedGradClipper()([p_.grad for p_ in params], 1.0)
RAdam(params, lr=1e-3).step()
# ---- kernels of round 2: GELU (odd row counts: the prefetch of the software pipeline runs past the last row), rotary
#      pair on column slices with column sums, embedding gradient (skewed ids), the fused feed-forward / q-k-v nodes
from sparse_vae_b200.core.attention import Attention  # noqa: E402
from sparse_vae_b200.core.embedding import Embedding  # noqa: E402
from sparse_vae_b200.core.gelu import GELU, ffn_forward, gelu_backward, gelu_forward  # noqa: E402
from sparse_vae_b200.core.linear import WeightShadows, rotary_pair  # noqa: E402
for rows, n in ((1, 8), (37, 24), (1001, 520), (4099, 2048)):
    xg = torch.randn(rows, n, device=dev).to(torch.bfloat16)
    gelu_forward(xg)
    gelu_backward(torch.randn_like(xg), xg, want_colsum=True)
    gelu_backward(torch.randn_like(xg), xg, inplace=True)
buf = torch.randn(3, 96, 3 * 520, device=dev).to(torch.bfloat16)
ang = torch.rand(96, 260, device=dev) * 6
rotary_pair(buf[..., :520], buf[..., 520:1040], ang.cos(), ang.sin(), conj=True, want_colsum=True, inplace=True)
rotary_pair(buf[..., :520], buf[..., 520:1040], ang.cos(), ang.sin())
emb = Embedding(1000, 64).to(dev)
ids = torch.randint(0, 1000, (5, 333), device=dev)
ids[ids % 3 == 0] = 7
emb(ids).sum().backward()
ffn = torch.nn.Sequential(Linear(512, 2048), GELU(), Linear(2048, 512, bias=False)).to(dev)
att = Attention(512, 8, causal=True, sparse=4).to(dev)
shadows = WeightShadows(torch.nn.ModuleList([ffn, att]))
xa = torch.randn(2, 1024, 512, device=dev, requires_grad=True)
with torch.autocast('cuda', dtype=torch.bfloat16), shadows.step():
    y = ffn_forward(ffn, xa) + att(xa, xa, xa)
y.float().sum().backward()
torch.cuda.synchronize()
print('sanitize_small: all kernels ran')
