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
torch.cuda.synchronize()
print('sanitize_small: all kernels ran')
