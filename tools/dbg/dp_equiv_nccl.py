"""2 ranks (NCCL) vs the same process's own full-batch gradients, per parameter."""
import os, sys
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / 'tests'))
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import test_gpu_data_parallel_nccl as T
from sparse_vae_b200.data_parallel import GradientAllReducer, init_distributed
rank, local, world = init_distributed('nccl')
dev = torch.device('cuda', local); torch.cuda.set_device(dev)
model = T._model(dev)
names = {id(p): n for n, p in model.named_parameters()}
lf = T._step(model, T._batch(dev, 0, T.B_GLOBAL)).item()
full = {names[id(p)]: p.grad.detach().clone() for p in model.parameters() if p.grad is not None}
reducer = GradientAllReducer(model, bucket_mb=4.0)
per = T.B_GLOBAL // world
for it in range(2):
    l = T._step(model, T._batch(dev, rank * per, (rank + 1) * per), reducer).item()
    red = {names[id(p)]: p.grad.detach().clone() for p in model.parameters() if p.grad is not None}
    if rank == 0:
        rows = sorted(((red[n] - g).norm().item() / (g.norm().item() + 1e-30), g.norm().item(), red[n].norm().item(), n) for n, g in full.items())[::-1]
        print(f'it {it}: loss full {lf} rank0 {l}; params {len(full)} vs {len(red)}; buckets {len(reducer.buckets)}')
        for r in rows[:12]: print(f'   {r[0]:.3e} |full| {r[1]:.4e} |reduced| {r[2]:.4e} {r[3]}')
dist.barrier(); dist.destroy_process_group()
