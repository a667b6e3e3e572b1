import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from sparse_vae_b200 import _native as N
from sparse_vae_b200.core.linear import rotary_pair
from sparse_vae_b200.core.attention import _cached_tables
B, L, d = 16, 4096, 512
a, b = (torch.randn(B, L, d, device='cuda').to(torch.bfloat16) for _ in range(2))
with torch.autocast('cuda', dtype=torch.bfloat16):
    cos, sin = _cached_tables(L, d // 2, 0, 256, torch.bfloat16, a.device)
def bench(fn, name, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps): fn()
    t1.record(); torch.cuda.synchronize(); print(f'{name:40s} {t0.elapsed_time(t1) / reps * 1e3:8.1f} us')
out = torch.empty_like(a)
def single(x, conj): N.check(N.lib.svae_rotary(x.data_ptr(), cos.data_ptr(), sin.data_ptr(), out.data_ptr(), N.DTYPE_BF16, N.svae_dtype(cos.dtype), B * L, L, d, conj, N.current_stream(x.device)), 'r')
bench(lambda: (single(a, 0), single(b, 0)), 'two single launches')
bench(lambda: rotary_pair(a, b, cos, sin), 'pair, forward')
bench(lambda: rotary_pair(a, b, cos, sin, conj=True, want_colsum=True), 'pair, backward + column sums')
