import torch
dev = torch.device('cuda'); M = 65536
def bench(fn, flops, name, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps): fn()
    t1.record(); torch.cuda.synchronize(); us = t0.elapsed_time(t1) / reps * 1e3
    print(f'{name:58s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s')
x = torch.randn(M, 512, device=dev, dtype=torch.bfloat16)
w = torch.randn(1536, 512, device=dev, dtype=torch.bfloat16) * 0.02; b = torch.randn(1536, device=dev, dtype=torch.bfloat16)
ws = [w[i * 512:(i + 1) * 512] for i in range(3)]; bs = [b[i * 512:(i + 1) * 512] for i in range(3)]
g = torch.randn(M, 1536, device=dev, dtype=torch.bfloat16)
gs = [g[:, i * 512:(i + 1) * 512] for i in range(3)]            # strided column slices
gc = [t.contiguous() for t in gs]
F = torch.nn.functional; S = 2 * M * 512 * 512
bench(lambda: [F.linear(x, ws[i], bs[i]) for i in range(3)], 3 * S, 'fwd: three [M,512]x[512,512]+b')
bench(lambda: F.linear(x, w, b), 3 * S, 'fwd: one [M,512]x[512,1536]+b')
def dgrad3():
    dx = torch.mm(gc[0], ws[0]); dx.addmm_(gc[1], ws[1]); dx.addmm_(gc[2], ws[2]); return dx
bench(dgrad3, 3 * S, 'dgrad: mm + 2 addmm_ (contiguous grads)')
bench(lambda: torch.mm(g, w), 3 * S, 'dgrad: one [M,1536]x[1536,512]')
bench(lambda: [torch.mm(gc[i].t(), x, out_dtype=torch.float32) for i in range(3)], 3 * S, 'wgrad: three [512,M]x[M,512] fp32 out')
bench(lambda: torch.mm(g.t(), x, out_dtype=torch.float32), 3 * S, 'wgrad: one [1536,M]x[M,512] fp32 out')
bench(lambda: [torch.mm(gs[i].t(), x, out_dtype=torch.float32) for i in range(3)], 3 * S, 'wgrad: three, strided column slices of [M,1536]')
bench(lambda: [torch.mm(gs[i], ws[i]) for i in range(3)], 3 * S, 'dgrad pieces on strided slices (no accumulate)')
