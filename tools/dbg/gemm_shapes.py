"""cuBLAS throughput at the GEMM shapes of one decoder block (bf16, fp32 accumulate) on this GPU."""
import torch
dev = torch.device('cuda')
M = 65536
def bench(fn, flops, name, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps): fn()
    t1.record(); torch.cuda.synchronize()
    us = t0.elapsed_time(t1) / reps * 1e3
    print(f'{name:46s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s')
x512 = torch.randn(M, 512, device=dev, dtype=torch.bfloat16)
x2048 = torch.randn(M, 2048, device=dev, dtype=torch.bfloat16)
w_up = torch.randn(2048, 512, device=dev, dtype=torch.bfloat16) * 0.02
b_up = torch.randn(2048, device=dev, dtype=torch.bfloat16)
w_dn = torch.randn(512, 2048, device=dev, dtype=torch.bfloat16) * 0.02
w_sq = torch.randn(512, 512, device=dev, dtype=torch.bfloat16) * 0.02
b_sq = torch.randn(512, device=dev, dtype=torch.bfloat16)
w_qkv = torch.randn(1536, 512, device=dev, dtype=torch.bfloat16) * 0.02
b_qkv = torch.randn(1536, device=dev, dtype=torch.bfloat16)
F = torch.nn.functional
G = 2 * M * 512 * 2048
bench(lambda: F.linear(x512, w_up, b_up), G, 'ffn up   fwd  [M,512]x[512,2048]+b')
bench(lambda: F.linear(x512, w_up), G, 'ffn up   fwd  no bias')
bench(lambda: F.linear(x2048, w_dn), G, 'ffn down fwd  [M,2048]x[2048,512]')
bench(lambda: x512 @ w_dn, G, 'ffn down dgrad [M,512]x[512,2048]')
bench(lambda: x2048 @ w_up, G, 'ffn up   dgrad [M,2048]x[2048,512]')
bench(lambda: torch.mm(x2048.t(), x512, out_dtype=torch.float32), G, 'ffn wgrad [2048,M]x[M,512] fp32 out')
bench(lambda: torch.mm(x512.t(), x2048, out_dtype=torch.float32), G, 'ffn wgrad [512,M]x[M,2048] fp32 out')
S = 2 * M * 512 * 512
bench(lambda: F.linear(x512, w_sq, b_sq), S, 'proj fwd [M,512]x[512,512]+b')
bench(lambda: F.linear(x512, w_qkv, b_qkv), 3 * S, 'qkv fwd as one [M,512]x[512,1536]+b')
bench(lambda: x512 @ w_sq, S, 'proj dgrad')
bench(lambda: torch.mm(x512.t(), x512, out_dtype=torch.float32), S, 'proj wgrad fp32 out')
y = torch.empty(M, 2048, device=dev, dtype=torch.bfloat16)
bench(lambda: F.gelu(x2048), 0, 'gelu fwd [M,2048]')
bench(lambda: y.copy_(x2048), 0, 'copy [M,2048] bf16 (537 MB)')
big = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
bench(lambda: big @ big, 2 * 8192 ** 3, '8192^3 (peak reference)')
