"""Data parallelism for the TransformerVAE step: one process per GPU, batch sharded across ranks, ONE collective --
the gradient all-reduce (mean) -- through torch.distributed (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in
the CPU tests).  The reference has no distributed code of its own; equal per-rank batches make
the averaged gradient equal to the global-batch gradient because both loss terms are batch means
(core/continuous_autoencoder.py:47, core/language_model.py:161-170).

`GradientAllReducer` keeps every grad-bearing parameter's `.grad` as a view into a few flat fp32 buckets and
launches each bucket's asynchronous all-reduce from a post-accumulate hook as soon as its last gradient has been
written, so communication overlaps the rest of backward.  Parameters that never receive a gradient (the
reference's unused `pos_linear` layers, core/attention.py:39) are discovered on the first step and left out.
Gradient clipping must run after `finish()`.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
from torch import nn


class _Bucket:
    def __init__(self, params: List[nn.Parameter], device, dtype):
        self.params = params
        self.numel = sum(p.numel() for p in params)
        self.flat = torch.zeros(self.numel, device=device, dtype=dtype)
        off = 0
        for p in params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.pending = len(params)
        self.work = None


class GradientAllReducer:
    def __init__(self, module: nn.Module, process_group: Optional[dist.ProcessGroup] = None, bucket_mb: float = 32.0):
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_bytes = int(bucket_mb * 2 ** 20)
        self.buckets: List[_Bucket] = []
        self._bucket_of = {}
        self._ready_order: List[nn.Parameter] = []
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad)
                         for p in module.parameters() if p.requires_grad]
        self._built = False

    # ---- hooks ----------------------------------------------------------------------------------
    def _on_grad(self, p: nn.Parameter):
        if not self._built:
            self._ready_order.append(p)
            return
        b = self._bucket_of.get(id(p))
        if b is None:
            return
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket):
        if self.world > 1:
            b.flat.mul_(1.0 / self.world)
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _build(self):
        """After the first backward: bucket the parameters that actually got gradients, in the order they became ready."""
        seen, order = set(), []
        for p in self._ready_order:
            if id(p) not in seen and p.grad is not None:
                seen.add(id(p))
                order.append(p)
        cur, cur_bytes = [], 0
        groups = []
        for p in order:
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= self.bucket_bytes:
                groups.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            groups.append(cur)
        for grp in groups:
            old = [p.grad.detach().clone() for p in grp]
            b = _Bucket(grp, grp[0].device, torch.float32 if grp[0].dtype != torch.float64 else torch.float64)
            for p, g in zip(grp, old):
                p.grad.copy_(g)
            for p in grp:
                self._bucket_of[id(p)] = b
            self.buckets.append(b)
        self._built = True
        self._ready_order = []

    # ---- per-step API ------------------------------------------------------------------------------
    def finish(self):
        """Call after backward(): waits for (or, on the first step, performs) the all-reduce of every bucket."""
        if not self._built:
            self._build()
            for b in self.buckets:
                self._launch(b)
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()
                b.work = None
            b.pending = len(b.params)

    def zero_grad(self):
        if not self._built:
            self.module.zero_grad(set_to_none=True)
            return
        for b in self.buckets:
            b.flat.zero_()

    @property
    def reduced_numel(self) -> int:
        return sum(b.numel for b in self.buckets)

    def remove(self):
        for h in self._handles:
            h.remove()


def init_distributed(backend: Optional[str] = None):
    """(rank, local_rank, world) from the torchrun environment; initialises the default process group if world > 1."""
    import os
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        kwargs = {}
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            kwargs['device_id'] = torch.device('cuda', local_rank)
        dist.init_process_group(backend=backend, **kwargs)
    return rank, local_rank, world
