"""Where does one training step go?  Top CUDA kernels by device time over 3 steps (torch.profiler; not a bench)."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import sparse_vae_b200 as sv  # noqa: E402
from sparse_vae_b200.core.lightning_shim import to_attrdict  # noqa: E402
from sparse_vae_b200.data_parallel import GradientAllReducer  # noqa: E402
from sparse_vae_b200.synthetic import synthetic_tokens, to_device  # noqa: E402

dev = torch.device('cuda')
B, L = 16, 4096
torch.manual_seed(7295)
model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams())).to(dev)
model.initialize_weights()
(opt,), _ = model.configure_optimizers(tokens_per_batch=B * L)
reducer = GradientAllReducer(model)
batch = to_device(synthetic_tokens(B, L), dev)


def step():
    reducer.zero_grad()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        loss = model.training_step(batch, 0)['loss']
    loss.backward()
    reducer.finish()
    model.on_after_backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=70))
# kernel-level breakdown (device events only), per step
from collections import defaultdict
agg = defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name][0] += 1
        agg[e.name][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f'KERNELS: {tot / 3e3:.2f} ms of device time per step')
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f'{us / 3e3:8.3f} ms/step {n // 3:5d} x {us / n:8.1f} us  {name[:160]}')
