"""One training step as ONE CUDA-graph replay (forward, backward, gradient all-reduce, clipping, RAdam).

A step of the default model is ~800 kernel launches of 10-60 us each; issued one by one from Python they leave ~3 ms
of gaps per 29 ms step and make the step time depend on how fast the host thread happens to run.  Captured once and
replayed, the launches cost the host nothing.  What changes from step to step is kept OUT of the captured launch
parameters and read from device memory that the host refreshes right before each replay:

* random numbers: the library's own kernels (bottleneck eps, fused dropout) take `{seed, base offset}` from a device
  buffer plus a per-launch increment fixed at capture time -- torch's own scheme for its generator
  (`offset_intragraph`); the default CUDA generator is advanced by the step's total.  ATen's own random ops inside the
  step (the samples of `marginal_kl`, ATen dropout where the fused one does not apply) are handled by torch's
  graph-safe generator registration and take the offsets AFTER the library's: a replayed step draws the library's
  numbers from [base, base + delta) and ATen's from there on -- the same offsets as the eager step whenever ATen's draws
  come last in program order (dropout off), disjoint ones otherwise.
* RAdam's step-dependent scalars (bias corrections, rectification, the scheduled learning rate): `svae_radam_args`
  evaluates them on the host in double precision exactly like the eager path and the kernel reads the block from
  device memory (`svae_radam_step_g`).
* the batch: copied into static input tensors.

With more than one process the NCCL calls stay OUTSIDE the graphs (a collective launched from an autograd hook inside a
capture deadlocked on this stack) but still overlap the backward: graph A = forward + backward with the reducer's hooks
live -- each gradient bucket is packed by a captured launch as soon as its last gradient exists, followed by an EXTERNAL
event-record node -- and right after `graph A.replay()` the host issues the bucket all-reduces on a side stream, each
behind its bucket's event (`GradientAllReducer.replay_reduce`); graph B = clipping + RAdam on the reduced buckets waits
for the last of them.  (`SVAE_DP_OVERLAP=0`: pack and reduce after graph A has finished, the first version: +1.7 ms.)

Reference call sites of what is captured: `TransformerVAE.training_step` (transformer_vae.py:42-66),
`LanguageModel.on_after_backward` (core/language_model.py:120-122), `RAdam.step` (core/rectified_adam.py:15-88).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import torch

from .. import _native as N


class StepPhilox:
    """Device-resident {seed, base offset} for the library's random kernels inside a captured step."""
    active: Optional['StepPhilox'] = None          # set while the step is being captured

    def __init__(self, device: torch.device):
        self.dev = torch.zeros(2, dtype=torch.int64, device=device)
        self.host = torch.zeros(2, dtype=torch.int64).pin_memory()
        self.delta = 0                             # running increment of the draws captured so far (multiple of 4)

    def reserve(self, increment: int):
        """(device pointer to {seed, base}, this draw's offset over the base)."""
        assert increment % 4 == 0
        d = self.delta
        self.delta += increment
        return self.dev.data_ptr(), d

    def refresh(self, generator: torch.Generator):
        """Before a replay: this step starts at the generator's current offset; the generator moves past the step."""
        seed, base = generator.initial_seed(), generator.get_offset()
        generator.set_offset(base + self.delta)
        self.host[0] = seed - (1 << 64) if seed >= (1 << 63) else seed
        self.host[1] = base
        self.dev.copy_(self.host, non_blocking=True)


class StepOptimArgs:
    """Device-resident scalar blocks of the fused RAdam step, one per parameter group."""
    active: Optional['StepOptimArgs'] = None

    def __init__(self, device: torch.device):
        self.device = device
        self.nbytes = int(N.lib.svae_radam_args_bytes())
        self.groups = {}                           # id(group) -> (group, host block, device block)

    def block_for(self, group: dict) -> int:
        ent = self.groups.get(id(group))
        if ent is None:
            host = torch.zeros(self.nbytes, dtype=torch.uint8).pin_memory()
            ent = self.groups[id(group)] = (group, host, torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device))
        return ent[2].data_ptr()

    def refresh(self):
        """Before a replay: the scalars of THIS step (group['lr'] as the scheduler left it, group['step'])."""
        for group, host, dev in self.groups.values():
            beta1, beta2 = group['betas']
            N.check(N.lib.svae_radam_args(float(group['lr']), float(beta1), float(beta2), float(group['eps']),
                                          float(group['weight_decay']), int(group.get('step', 1)), host.data_ptr()), 'svae_radam_args')
            dev.copy_(host, non_blocking=True)

    def advance(self):
        for group, _, _ in self.groups.values():
            group['step'] = group.get('step', 1) + 1


class GraphedTrainStep:
    """`loss = step(batch)` with the same effect as

        reducer.zero_grad(); loss = model.training_step(batch, 0)['loss'] (under autocast); loss.backward();
        reducer.finish(); model.on_after_backward(); optimizer.step(); scheduler.step(); model.global_step += 1

    The first `warmup` calls run exactly that, eagerly (they also build the all-reduce buckets, the optimizer state and
    the 16-bit weight copies); the next call captures the device work of one step into a CUDA graph, and from then
    on a call is: refresh the device-resident step parameters, copy the batch into the static inputs, replay.
    The returned loss is a static tensor that the next replay overwrites.
    """

    def __init__(self, model, optimizer, scheduler=None, reducer=None, autocast_dtype=torch.bfloat16, warmup: int = 3):
        self.model, self.opt, self.sched, self.reducer = model, optimizer, scheduler, reducer
        self.autocast_dtype = autocast_dtype
        self.warmup = max(int(warmup), 1)              # the first step also discovers grad-less parameters
        if getattr(model, 'validate_posterior', False) is None:
            # torch's default: Normal(...) checks loc / scale on the HOST, a device synchronisation that cannot be captured
            model.validate_posterior = False
        self.calls = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_b: Optional[torch.cuda.CUDAGraph] = None          # world > 1: clipping + optimizer, after the all-reduce
        self.split = reducer is not None and getattr(reducer, 'world', 1) > 1
        # world > 1: all-reduce each bucket while the replayed backward is still running (SVAE_DP_OVERLAP=0: after it)
        self.overlap = self.split and os.environ.get('SVAE_DP_OVERLAP', '1') != '0'
        self._captured_grads = None
        self.static_batch: Optional[Dict[str, torch.Tensor]] = None
        self.static_loss: Optional[torch.Tensor] = None
        self.philox: Optional[StepPhilox] = None
        self.optim_args: Optional[StepOptimArgs] = None

    # ---- the step, as the eager trainer runs it ---------------------------------------------------------------
    def _device_work(self, batch):
        if self.reducer is not None:
            self.reducer.zero_grad()
        else:
            self.model.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=self.autocast_dtype):
            out = self.model.training_step(batch, 0)
        out['loss'].backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.model.on_after_backward()                 # gradient clipping (after the all-reduce)
        self.opt.step()
        return out['loss'].detach()

    def _host_bookkeeping(self):
        if self.sched is not None:
            self.sched.step()
        self.model.global_step += 1

    def eager(self, batch):
        """The same step, launch by launch (warm-up, and per-kernel profiling: a replay runs no host code to time)."""
        loss = self._device_work(batch)
        self._host_bookkeeping()
        return loss

    # ---- capture / replay ---------------------------------------------------------------------------------------
    def _raw(self, batch):
        tok = batch['token_ids']
        tok = tok.as_raw() if hasattr(tok, 'as_raw') else tok
        return tok, batch['num_tokens']

    def _capture(self, batch):
        from .padded_tensor import PaddedTensor
        hp = self.model.hparams
        if getattr(hp, 'kl_annealing_steps', 0):
            raise RuntimeError("kl_weight annealing changes a captured constant every step; train eagerly or anneal first")
        dev = next(self.model.parameters()).device
        tok, n = self._raw(batch)
        s_tok, s_n = tok.clone(), n.clone()
        self.static_batch = {'token_ids': PaddedTensor.from_raw(s_tok), 'num_tokens': s_n, 'num_bytes': 4 * s_n}
        self._static_raw = (s_tok, s_n)
        self.philox, self.optim_args = StepPhilox(dev), StepOptimArgs(dev)
        for g in self.opt.param_groups:                # allocated NOW: a tensor created inside the capture belongs to the
            self.optim_args.block_for(g)               # graph's pool and its zero-fill would be replayed over the refresh
        self.graph = torch.cuda.CUDAGraph()
        steps_before = [g.get('step', 1) for g in self.opt.param_groups]
        torch.cuda.synchronize(dev)
        StepPhilox.active, StepOptimArgs.active = self.philox, self.optim_args
        try:
            if not self.split:
                with torch.cuda.graph(self.graph):
                    self.static_loss = self._device_work(self.static_batch)
            else:
                if self.overlap:
                    # hooks live: every bucket is packed inside the graph as soon as its last gradient exists and an
                    # external event marks the spot; the collectives themselves are issued eagerly after the replay
                    try:
                        with torch.cuda.graph(self.graph):
                            self.reducer.zero_grad()
                            self.reducer.begin_capture()
                            with torch.autocast('cuda', dtype=self.autocast_dtype):
                                out = self.model.training_step(self.static_batch, 0)
                            out['loss'].backward()
                            self.reducer.end_capture()     # `.grad` -> bucket views, which graph B reads
                            self.static_loss = out['loss'].detach()
                    except Exception as exc:               # e.g. a torch build without external events: reduce after graph A
                        import warnings
                        warnings.warn(f"overlapped gradient all-reduce could not be captured ({type(exc).__name__}: {exc}); "
                                      f"reducing after the backward graph instead", RuntimeWarning)
                        self.reducer._capturing = False
                        self.overlap = False
                        self.graph = torch.cuda.CUDAGraph()
                        self.philox.delta = 0
                        torch.cuda.synchronize(dev)
                if not self.overlap:
                    self.reducer.sync = False          # the hooks stay quiet: pack + reduce after the replay
                    try:
                        with torch.cuda.graph(self.graph):
                            self.reducer.zero_grad()
                            with torch.autocast('cuda', dtype=self.autocast_dtype):
                                out = self.model.training_step(self.static_batch, 0)
                            out['loss'].backward()
                            self.static_loss = out['loss'].detach()
                    finally:
                        self.reducer.sync = True
                    self._captured_grads = {id(p): p.grad for p in self.model.parameters() if p.grad is not None}
                    self.reducer.reduce_tensors(self._captured_grads)    # `.grad` -> bucket views, which graph B reads
                self.graph_b = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_b, pool=self.graph.pool()):
                    self.model.on_after_backward()
                    self.opt.step()
        finally:
            StepPhilox.active = StepOptimArgs.active = None
        # capturing ran the optimizer's Python (which counts a step) without executing anything on the device
        for g, st in zip(self.opt.param_groups, steps_before):
            g['step'] = st

    def __call__(self, batch):
        self.calls += 1
        if self.calls <= self.warmup:
            return self.eager(batch)
        if self.graph is None:
            self._capture(batch)
        tok, n = self._raw(batch)
        s_tok, s_n = self._static_raw
        if tok.data_ptr() != s_tok.data_ptr():
            s_tok.copy_(tok, non_blocking=True)
            s_n.copy_(n, non_blocking=True)
            self.static_batch['num_bytes'].copy_(4 * s_n)
        dev = s_tok.device
        gen = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
        self.philox.refresh(gen)
        self.optim_args.refresh()
        self.graph.replay()
        if self.split:
            if self.overlap:
                self.reducer.replay_reduce()
            else:
                self.reducer.reduce_tensors(self._captured_grads)
            self.graph_b.replay()
        self.optim_args.advance()
        self._host_bookkeeping()
        return self.static_loss
