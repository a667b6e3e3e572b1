"""Data parallelism for the TransformerVAE step: one process per GPU, batch sharded across ranks, ONE collective --
the gradient all-reduce (mean) -- through torch.distributed (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in
the CPU tests).  The reference has no distributed code of its own; equal per-rank batches make
the averaged gradient equal to the global-batch gradient because both loss terms are batch means
(core/continuous_autoencoder.py:47, core/language_model.py:161-170).

`GradientAllReducer` packs the gradients of every grad-bearing parameter into a few flat fp32 buckets (one
multi-tensor launch per bucket, with the 1/world averaging folded in) from a post-accumulate hook as soon as the
bucket's last gradient has been written, and launches the bucket's asynchronous all-reduce at once, so
communication overlaps the rest of backward; afterwards every `.grad` is a view into its bucket.  With one process
there is nothing to reduce and the gradients stay where autograd put them.  Parameters that never receive a gradient (the
reference's unused `pos_linear` layers, core/attention.py:39) are discovered on the first step and left out.
Gradient clipping must run after `finish()`.  With gradient accumulation set `reducer.sync = False` for every
micro-batch but the last one (the hooks fire on every backward).
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
        self.views, off = [], 0
        for p in params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.pending = len(params)
        self.work = None
        self._fused = None
        self.event = None           # external event recorded inside a captured backward once the bucket is packed

    def gather(self, scale: float):
        """Packs the gradients autograd left on the parameters into the flat buffer (times `scale`) and re-points
        every `.grad` at its slice: one multi-tensor launch on CUDA (csrc/optim.cu), plain copies elsewhere."""
        todo = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        stale = [v for v, p in zip(self.views, self.params) if p.grad is None]
        # a gradient that already lives in its bucket slice (the step started with gradients kept, e.g.
        # zero_grad(set_to_none=False): autograd accumulated straight into the view) still needs the 1/world factor
        aliased = [v for v, p in zip(self.views, self.params) if p.grad is not None and p.grad.data_ptr() == v.data_ptr()]
        if aliased and scale != 1.0:
            torch._foreach_mul_(aliased, scale)
        if todo:
            dst, src = [v for v, _ in todo], [g for _, g in todo]
            fused = self.flat.is_cuda and all(g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() for g in src) \
                and self.flat.dtype == torch.float32
            if fused:
                if self._fused is None:
                    from .fused_optim import FusedScaleCopy
                    self._fused = FusedScaleCopy()
                self._fused(dst, src, scale)
            else:
                for v, g in todo:
                    v.copy_(g)
                    if scale != 1.0:
                        v.mul_(scale)
        for v in stale:                       # a parameter that got no gradient this step contributes zeros
            v.zero_()
        for v, p in zip(self.views, self.params):
            p.grad = v


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
        self._reduced_numel = 0
        self.sync = True        # set False for all but the last micro-batch of a gradient-accumulation step
        self._capturing = False  # inside a CUDA-graph capture of forward + backward (begin_capture / end_capture)
        self._capture_order: List[_Bucket] = []
        self._comm_stream = None

    # ---- hooks ----------------------------------------------------------------------------------
    def _on_grad(self, p: nn.Parameter):
        if not self.sync:                     # gradients keep accumulating locally; reduce on the last micro-batch
            return
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
        # gradients arrive as separate tensors (autograd assigns them: `.grad` is None at the start of a step, so
        # nothing is zero-filled or accumulated); one launch packs the bucket and applies the 1/world averaging
        b.gather(1.0 / self.world)
        if self._capturing:
            # the pack is part of the graph; the collective is NOT (see begin_capture): an event-record node marks the
            # point of the replayed backward from which the bucket may be reduced
            if b.event is None:
                b.event = torch.cuda.Event(external=True)
            b.event.record()
            b.work = 'captured'
            self._capture_order.append(b)
            return
        b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _build(self):
        """After the first backward: bucket the parameters that actually got gradients, in the order they became ready."""
        seen, order = set(), []
        for p in self._ready_order:
            if id(p) not in seen and p.grad is not None:
                seen.add(id(p))
                order.append(p)
        self._reduced_numel = sum(p.numel() for p in order)
        self._built = True
        self._ready_order = []
        if self.world == 1:                   # nothing to reduce: gradients stay where autograd put them
            return
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
            b = _Bucket(grp, grp[0].device, torch.float32 if grp[0].dtype != torch.float64 else torch.float64)
            for p in grp:
                self._bucket_of[id(p)] = b
            self.buckets.append(b)

    # ---- per-step API ------------------------------------------------------------------------------
    def finish(self):
        """Call after backward(): waits for (or, on the first step, performs) the all-reduce of every bucket."""
        if not self._built:
            self._build()
            for b in self.buckets:
                self._launch(b)
        for b in self.buckets:
            if b.work is None and self.sync:
                # some (or all) parameters of the bucket got no gradient on THIS rank this step: their slices are
                # zero-filled and the bucket is reduced anyway, so that every rank issues the same collectives
                self._launch(b)
            if b.work is not None:
                b.work.wait()
                b.work = None
            b.pending = len(b.params)

    # ---- captured backward: the all-reduce overlaps the replayed graph ---------------------------------------------
    def begin_capture(self):
        """Call inside `torch.cuda.graph(...)` before the forward pass.  The hooks stay live during the captured
        backward: each bucket is packed by a captured launch as soon as its last gradient exists, followed by an
        EXTERNAL event-record node.  No collective is captured (an NCCL call issued from the autograd thread inside a
        capture deadlocked on this stack); `replay_reduce()` issues them eagerly on a side stream, each behind its
        bucket's event, so they run while the rest of the replayed backward is still executing."""
        assert self._built and self.world > 1, "run one eager step first: the buckets are laid out after the first backward"
        self._capturing = True
        self._capture_order = []
        for b in self.buckets:
            b.pending, b.work = len(b.params), None

    def end_capture(self):
        """Call inside the capture after backward(): buckets whose hook never completed (a parameter without a
        gradient on this step) are packed here, zero-filled where nothing arrived."""
        for b in self.buckets:
            if b.work is None:
                self._launch(b)
        self._capturing = False
        for b in self.buckets:
            b.pending, b.work = len(b.params), None

    def replay_reduce(self):
        """Right after `graph.replay()` of a backward captured between begin_capture / end_capture: all-reduce every
        bucket on the communication stream as soon as the replay has packed it; the caller's stream then waits for all
        of them (clipping and the optimizer read the bucket views, which the `.grad` attributes already are)."""
        main = torch.cuda.current_stream()
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream()
        works = []
        with torch.cuda.stream(self._comm_stream):
            for b in self._capture_order:
                self._comm_stream.wait_event(b.event)
                works.append(dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()                          # the caller's stream waits for the collective (no host block)
        main.wait_stream(self._comm_stream)

    def reduce_tensors(self, grads_by_param: dict):
        """All-reduce (mean) of gradients that live OUTSIDE the `.grad` attributes -- the static gradient tensors of a
        captured backward (core/graph_step.py): packed into the flat buckets (1/world folded in), reduced, waited for;
        afterwards every `.grad` is the bucket view.  `grads_by_param`: {id(parameter): gradient tensor}."""
        assert self._built, "run one eager step first: the buckets are laid out after the first backward"
        for b in self.buckets:
            dst, src = [], []
            for v, p in zip(b.views, b.params):
                g = grads_by_param.get(id(p))
                if g is None:
                    v.zero_()
                else:
                    dst.append(v)
                    src.append(g)
            if dst:
                if b.flat.is_cuda and all(g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() for g in src):
                    if b._fused is None:
                        from .fused_optim import FusedScaleCopy
                        b._fused = FusedScaleCopy()
                    b._fused(dst, src, 1.0 / self.world)
                else:
                    for v, g in zip(dst, src):
                        v.copy_(g)
                        v.mul_(1.0 / self.world)
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        for b in self.buckets:
            b.work.wait()
            b.work = None
            b.pending = len(b.params)
            for v, p in zip(b.views, b.params):
                p.grad = v

    def zero_grad(self):
        self.module.zero_grad(set_to_none=True)

    @property
    def reduced_numel(self) -> int:
        return self._reduced_numel

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
