#!/usr/bin/env python
"""`python train.py transformer-vae [preset=NAME] [model.key=value ...] [trainer.max_steps=N]`

Command-line surface of the reference's train.py (model name, `key=value` dotlist, named preset merged last, seed
7295, default `accumulate_grad_batches=2`) driving this package's TransformerVAE on SYNTHETIC token batches: the
reference's Lightning trainer and HuggingFace data module are outside this build's scope and
there is no network for datasets.  Multi-GPU: launch with torchrun; gradients are all-reduced over NCCL.
"""
from __future__ import annotations

import sys

import torch

import sparse_vae_b200 as sv
from sparse_vae_b200.core.lightning_shim import AttrDict, to_attrdict
from sparse_vae_b200.data_parallel import GradientAllReducer, init_distributed
from sparse_vae_b200.synthetic import synthetic_tokens, to_device


def _parse_value(text: str):
    for cast in (int, float):
        try:
            return cast(text)
        except ValueError:
            pass
    return {'true': True, 'false': False, 'none': None, 'null': None}.get(text.lower(), text)


def main(args):
    if len(args) < 2 or args[1] != 'transformer-vae':
        print(f"Unrecognized model type '{args[1] if len(args) > 1 else ''}'. This build provides 'transformer-vae'.")
        sys.exit(1)
    torch.manual_seed(7295)
    config = AttrDict(trainer=AttrDict(accumulate_grad_batches=2, precision='bf16', max_steps=20),
                      model=to_attrdict(sv.TransformerVAEHparams()),
                      data=AttrDict(tokens_per_batch=50_000, seq_len=4096))
    for item in args[2:]:
        key, _, value = item.partition('=')
        node, *rest = key.split('.')
        if rest:
            config[node][rest[0]] = _parse_value(value)
        else:
            config[node] = _parse_value(value)
    if preset := config.get('preset'):
        chosen = sv.hparam_presets.get(preset)
        assert chosen, f"Preset name '{preset}' not recognized."
        for section, values in chosen.items():          # the preset wins over the command line, as in the reference
            config.setdefault(section, AttrDict()).update(values)

    rank, local_rank, world = init_distributed()
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    model = sv.TransformerVAE(to_attrdict({k: v for k, v in config.model.items()
                                           if k in sv.TransformerVAEHparams.__dataclass_fields__})).to(device)
    model.on_fit_start()
    model.on_train_start()
    seq_len = config.data.seq_len
    batch_size = max(1, config.data.tokens_per_batch // seq_len // world)
    accum = config.trainer.accumulate_grad_batches
    (opt,), (sched,) = model.configure_optimizers(tokens_per_batch=batch_size * seq_len * world,
                                                  accumulate_grad_batches=accum)
    reducer = GradientAllReducer(model)
    print(f"Training transformer-vae on synthetic tokens: {world} GPU(s), {batch_size} x {seq_len} per GPU per micro-batch")
    for step in range(config.trainer.max_steps):
        reducer.zero_grad()
        for micro in range(accum):
            reducer.sync = micro == accum - 1            # all-reduce only once the last micro-batch has been accumulated
            batch = to_device(synthetic_tokens(batch_size, seq_len, seed=7295 + 131 * step + micro + 7919 * rank), device)
            with torch.autocast('cuda', dtype=torch.bfloat16):
                loss = model.training_step(batch, step)['loss'] / accum
            loss.backward()
        reducer.finish()
        model.on_after_backward()
        opt.step()
        sched['scheduler'].step()
        model.global_step += 1
        if rank == 0:
            logged = {k: round(float(v), 4) for k, v in model.logged.items()}
            print(f"step {step}: loss {float(loss.detach()) * accum:.4f} {logged}")


if __name__ == '__main__':
    main(sys.argv)
