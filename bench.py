#!/usr/bin/env python
"""Benchmark of the sparse-vae hot path on B200: TransformerVAE training tokens/sec at seq 4096.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--kernel-only]

Workload (BASELINE.json configs[1], "C2"): TransformerVAE with the default hparams dataclass (d_model 512, 8 heads,
6 decoder layers + Perceiver encoder, block-sparse attention window 4, latent 64, tied 32768-token embedding,
grad_checkpointing off), batch 16 x 4096 synthetic tokens PER GPU, bf16 autocast, RAdam, gradient clipping; a step
is forward + backward + gradient all-reduce (N > 1) + clip + optimizer step.  N > 1 is launched by torchrun, one
rank per GPU, NCCL; weak scaling (per-GPU batch fixed).

One JSON line on stdout (rank 0):
  value        whole-job tokens/sec, inputs resident in HBM, K steps timed with CUDA events, max over ranks
  e2e          the same through the public API with HOST batches: every step copies the int16 token ids from pinned
               host memory and copies the loss back to pinned host memory (consumed by the host one step later)
  roofline     the dominant kernel of this library inside the timed region: algorithmic bytes per launch
                / mean launch duration (CUDA events on the launching stream, svae_profile_*)
  cpu_baseline the oracle's CPU restatement of the same training step on the host cores (bounded sample)
`--impl reference` times that CPU restatement alone (the reference itself is Python/Triton-1.1 and cannot run on
this box).  `--kernel-only` runs just the attention kernels at the C2 shape (short; for ncu).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = 'train tokens/sec at seq 4096'
UNIT = 'tokens/s'
SEQ, BATCH_PER_GPU = 4096, 16
H, DH = 8, 64
FALLBACK_HBM_GBS = 6650.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH_PER_GPU, help='sequences per GPU')
    ap.add_argument('--seq', type=int, default=SEQ)
    ap.add_argument('--kernel-only', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample-seqs', type=int, default=1)
    return ap.parse_args()


def workload_config(args, world):
    return {
        'workload': f'TransformerVAE default hparams (d_model 512, 8 heads, 6 layers, sparse window 4, latent 64), '
                    f'batch {args.batch} x seq {args.seq} per GPU, bf16 autocast, RAdam + grad clip',
        'global_batch': args.batch * world, 'seq_len': args.seq, 'parallelism': f'dp{world}',
        'accumulate_grad_batches': 1, 'grad_checkpointing': False,
        'host_syncs': 'none inside a step (validate_args=False on the returned posterior Normal; the reference checks it on the host)',
        'l2': 'no explicit flush: one step streams >10 GB of activations/weights/gradients through the 126 MB L2',
        'e2e_loss_read': 'every step: async D2H of the loss into pinned memory + event, read by the host one step later',
    }


def peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        d = json.loads(p.read_text())
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            parts = [x.strip() for x in s.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_step_fn(args, seqs: int):
    """The oracle's plain-torch restatement of TransformerVAE.training_step + backward, fp32, all host threads."""
    from oracle import model as omodel              # nothing of sparse_vae_b200 is on this path

    d_model, heads, layers, window, latent = 512, 8, 6, 4, 64       # TransformerVAEHparams() defaults
    params = omodel.init_params(d_model, layers, latent, seed=7295)
    g = torch.Generator().manual_seed(7295)
    tok = torch.randint(3, 2 ** 15, (seqs, args.seq), generator=g)
    tok[:, 0], tok[:, -1] = 1, 2
    counts = torch.full((seqs,), args.seq)

    def step():
        eps = torch.randn(seqs, 1, latent)
        out = omodel.training_step(params, tok, counts, eps, d_model, heads, layers, window, dropout_p=0.1)
        out['loss'].backward()
        for p in params.values():
            p.grad = None
        return float(out['loss'].detach())

    return step


def run_cpu_reference(args, steps: int, warmup: int, seqs: int):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_reference_step_fn(args, seqs)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {'value': seqs * args.seq / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'{seqs} sequence(s) x {args.seq} tokens per step (1/{max(args.batch // seqs, 1)} of the per-GPU batch), '
                      f'fp32, oracle/model.py training_step + backward, {steps} step(s) after {warmup} warm-up',
            'ms_per_step': dt * 1e3}


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(0, min(args.warmup, 1))
    res = run_cpu_reference(args, steps, warmup, args.cpu_sample_seqs)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': res['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warmup, 'ms_per_step': res['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args, 1),
        'cpu_baseline': {k: res[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': res['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ kernel figures
def algorithmic_bytes(kernel: str, B: int, L: int) -> float:
    """Algorithmic HBM bytes of ONE launch (DESIGN.md section "Kernels"): unit = one [B,L,H,Dh] 16-bit tensor."""
    unit = B * L * H * DH * 2
    stats = B * H * L * 4
    return {
        'attn_fwd_sm100': 4 * unit + stats,                  # read Q,K,V ; write O, LSE
        'attn_bwd_dq_sm100': 6 * unit + 2 * stats,           # read Q,K,V,O,dO,LSE ; write dQ, delta
        'attn_bwd_dkv_sm100': 6 * unit + 2 * stats,          # read K,V,Q,dO,LSE,delta ; write dK,dV
    }.get(kernel, 0.0)


def roofline_block(prof: dict, B: int, L: int):
    mine = {k: v for k, v in prof.items() if v['launches'] > 0}
    if not mine:
        return None, {}
    peak, which = peaks()
    per_kernel = {}
    for name, v in mine.items():
        us = v['ms'] / v['launches'] * 1e3
        ab = algorithmic_bytes(name, B, L)
        per_kernel[name] = {'launches': v['launches'], 'us_per_launch': us,
                            'algorithmic_gbs': (ab / (us * 1e-6) / 1e9) if ab else None}
    top = max((k for k in mine if algorithmic_bytes(k, B, L) > 0), key=lambda k: mine[k]['ms'], default=None)
    if top is None:
        return None, per_kernel
    achieved = per_kernel[top]['algorithmic_gbs']
    traffic = None
    tfile = ROOT / 'profiles' / 'ncu_traffic.json'
    if tfile.exists():
        traffic = json.loads(tfile.read_text()).get(top)
    return {'kernel': top, 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
            'traffic': traffic, 'peak_source': which,
            'algorithmic_bytes_per_launch': algorithmic_bytes(top, B, L)}, per_kernel


def kernel_only(args):
    """Attention forward + backward at the C2 shape, L2 flushed between iterations; per-kernel roofline."""
    import sparse_vae_b200 as sv
    from sparse_vae_b200 import _native as N
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    B, L = args.batch, args.seq
    cfg = sv.SparseAttention()
    g = torch.Generator().manual_seed(7295)
    q, k, v, do = (torch.randn(B, L, H * DH, generator=g).to(dev, torch.bfloat16).unflatten(-1, (H, DH)).transpose(1, 2)
                   for _ in range(4))
    q, k, v = (t.requires_grad_(True) for t in (q, k, v))
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)

    def it():
        flush.zero_()
        out = cfg(q, k, v)
        flush.zero_()
        out.backward(do)
        q.grad = k.grad = v.grad = None

    for _ in range(args.warmup):
        it()
    torch.cuda.synchronize()
    N.profile_begin()
    for _ in range(args.steps):
        it()
    torch.cuda.synchronize()
    prof = N.profile_end()
    roof, per_kernel = roofline_block(prof, B, L)
    print(json.dumps({'mode': 'kernel-only', 'shape': [B, H, L, DH], 'roofline': roof, 'kernels': per_kernel,
                      'l2': 'flushed (256 MiB memset) before every launch'}))


# ------------------------------------------------------------------------------------------ our arm
def main_ours(args):
    import torch.distributed as dist
    import sparse_vae_b200 as sv
    from sparse_vae_b200 import _native as N
    from sparse_vae_b200.core.lightning_shim import to_attrdict
    from sparse_vae_b200.data_parallel import GradientAllReducer, init_distributed
    from sparse_vae_b200.synthetic import synthetic_tokens, to_device

    rank, local_rank, world = init_distributed('nccl')
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    N.check(N.lib.svae_device_check(), 'svae_device_check')
    B, L = args.batch, args.seq

    torch.manual_seed(7295)                                  # same init on every rank
    hp = to_attrdict(sv.TransformerVAEHparams())
    model = sv.TransformerVAE(hp).to(dev)
    model.initialize_weights()
    model.train()
    model.validate_posterior = False       # no host-side check of the returned posterior: keeps the step free of device syncs
    (opt,), (sched_cfg,) = model.configure_optimizers(tokens_per_batch=B * L * world, accumulate_grad_batches=1)
    sched = sched_cfg['scheduler']
    reducer = GradientAllReducer(model)

    n_host = 4
    host = [synthetic_tokens(B, L, seed=7295 + rank * 1000 + i, pin=True) for i in range(n_host)]
    resident = [to_device(hb, dev, non_blocking=False) for hb in host]
    torch.manual_seed(7295 + rank)                           # per-rank dropout / eps streams

    def step(batch):
        reducer.zero_grad()
        with torch.autocast('cuda', dtype=torch.bfloat16):
            out = model.training_step(batch, 0)
        out['loss'].backward()
        reducer.finish()
        model.on_after_backward()                            # gradient clipping (after the all-reduce)
        opt.step()
        sched.step()
        model.global_step += 1
        return out['loss']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(steps):
            fn(i)
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for i in range(args.warmup):
        step(resident[i % n_host])

    # ---- device-resident inputs: the headline `value`
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    N.profile_begin()
    ms = timed(lambda i: step(resident[i % n_host]), args.steps)
    prof = N.profile_end()
    clocks = sampler.stop() if rank == 0 else None
    tokens = world * B * L * args.steps
    value = tokens / (ms * 1e-3)

    # ---- end to end through the public API with host batches (pinned H2D of the ids, D2H of the loss, every step)
    h2d = host[0]['token_ids'].numel() * host[0]['token_ids'].element_size() + host[0]['num_tokens'].numel() * 8
    losses = []

    # every step's loss goes device -> pinned host memory inside the timed region; the host consumes it one step later
    # (asynchronous logging), so the read does not drain the GPU at every step boundary
    loss_host = torch.empty(args.steps, dtype=torch.float32).pin_memory()
    loss_ready = [torch.cuda.Event() for _ in range(args.steps)]

    def e2e_step(i):
        batch = to_device(host[i % n_host], dev, non_blocking=True)
        loss_host[i:i + 1].copy_(step(batch).detach().reshape(1), non_blocking=True)
        loss_ready[i].record()
        if i > 0:
            loss_ready[i - 1].synchronize()
            losses.append(float(loss_host[i - 1]))

    ms_e2e = timed(e2e_step, args.steps)
    losses.append(float(loss_host[args.steps - 1]))          # timed() ended with a device synchronisation
    e2e_value = tokens / (ms_e2e * 1e-3)

    if rank == 0:
        roof, per_kernel = roofline_block(prof, B, L)
        launches = sum(v['launches'] for v in prof.values())
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = run_cpu_reference(args, steps=1, warmup=1, seqs=args.cpu_sample_seqs)
            cpu = {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16', 'data': 'synthetic', 'config': workload_config(args, world),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches, 'roofline': roof, 'kernels': per_kernel, 'cpu_baseline': cpu, 'clocks': clocks,
            'final_loss': losses[-1] if losses else None,
            'grad_allreduce_numel': reducer.reduced_numel,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        return main_reference(args)
    if args.kernel_only:
        return kernel_only(args)
    return main_ours(args)


if __name__ == '__main__':
    main()
