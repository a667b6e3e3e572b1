"""Which parameters' gradients differ between one process with the whole batch and the mean over two half batches?
(single GPU, no NCCL: the arithmetic identity the 2-rank test relies on)"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / 'tests'))
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import test_gpu_data_parallel_nccl as T

dev = torch.device('cuda', 0)
model = T._model(dev)
names = {id(p): n for n, p in model.named_parameters()}
def grads(lo, hi):
    loss = T._step(model, T._batch(dev, lo, hi)).item()
    return loss, {names[id(p)]: p.grad.detach().clone() for p in model.parameters() if p.grad is not None}
lf, full = grads(0, T.B_GLOBAL)
l0, g0 = grads(0, T.B_GLOBAL // 2)
l1, g1 = grads(T.B_GLOBAL // 2, T.B_GLOBAL)
print('loss full', lf, 'mean of halves', 0.5 * (l0 + l1))
rows = []
for n, g in full.items():
    m = 0.5 * (g0[n] + g1[n])
    rows.append(((m - g).norm().item() / (g.norm().item() + 1e-30), g.norm().item(), m.norm().item(), n))
rows.sort(reverse=True)
for r in rows[:25]:
    print(f'{r[0]:.3e}  |full| {r[1]:.4e}  |mean halves| {r[2]:.4e}  {r[3]}')
tot_f = torch.cat([g.flatten() for g in full.values()]).norm().item()
tot_m = torch.cat([(0.5 * (g0[n] + g1[n])).flatten() for n in full]).norm().item()
print('total norm', tot_f, tot_m)
