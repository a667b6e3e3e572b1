"""CPU, world_size 2, gloo: the gradient all-reduce used for data parallelism (sparse_vae_b200/data_parallel.py).

The product attention kernels are CUDA-only, so the host-side collective logic is exercised with a small dense
module that has, like the reference's Attention (`pos_linear`, core/attention.py:39), a parameter that never
receives a gradient.  Averaged per-rank gradients must equal the gradient of the global batch, step after step,
and be bitwise identical across ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(8, 16)
        self.unused = nn.Linear(8, 8)          # never touched in forward
        self.b = nn.Linear(16, 4)

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from sparse_vae_b200.data_parallel import GradientAllReducer, init_distributed
    r, _, w = init_distributed('gloo')
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    model = Toy()
    ref = Toy()
    ref.load_state_dict(model.state_dict())
    reducer = GradientAllReducer(model, bucket_mb=0.0001)          # tiny buckets -> several all-reduces
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    opt_ref = torch.optim.SGD(ref.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(1)
    ok = True
    for step in range(6):
        x = torch.randn(world * 6, 8, generator=g)
        y = torch.randn(world * 6, 4, generator=g)
        xs, ys = x[rank * 6:(rank + 1) * 6], y[rank * 6:(rank + 1) * 6]
        if step >= 4:       # gradients kept and zeroed in place: autograd accumulates straight into the bucket views
            model.zero_grad(set_to_none=False)
        else:
            reducer.zero_grad()
        if step % 2:                              # gradient accumulation over two micro-batches: reduce after the last one
            half = xs.shape[0] // 2
            reducer.sync = False
            (((model(xs[:half]) - ys[:half]) ** 2).mean() / 2).backward()
            reducer.sync = True
            (((model(xs[half:]) - ys[half:]) ** 2).mean() / 2).backward()
        else:
            ((model(xs) - ys) ** 2).mean().backward()
        reducer.finish()
        ref.zero_grad(set_to_none=True)
        ((ref(x) - y) ** 2).mean().backward()                       # global batch on one process
        for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            if n.startswith('unused'):
                ok &= p.grad is None
            else:
                ok &= torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-7)
        flat = torch.cat([p.grad.flatten() for n, p in model.named_parameters() if p.grad is not None])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        ok &= all(torch.equal(gathered[0], t) for t in gathered)    # bitwise identical on every rank
        opt.step()
        opt_ref.step()
    ok &= len(reducer.buckets) > 1 and reducer.reduced_numel == sum(p.numel() for n, p in model.named_parameters()
                                                                    if not n.startswith('unused'))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
