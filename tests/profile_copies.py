"""Which ops launch copy kernels in a training step?  (torch.profiler with shapes; debugging aid, not a bench)"""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import sparse_vae_b200 as sv  # noqa: E402
from sparse_vae_b200.core.lightning_shim import to_attrdict  # noqa: E402
from sparse_vae_b200.synthetic import synthetic_tokens, to_device  # noqa: E402

dev = torch.device('cuda')
torch.manual_seed(7295)
model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams())).to(dev)
model.initialize_weights()
(opt,), _ = model.configure_optimizers(tokens_per_batch=16 * 4096)
batch = to_device(synthetic_tokens(16, 4096), dev)


def step():
    for p in model.parameters():
        p.grad = None
    with torch.autocast('cuda', dtype=torch.bfloat16):
        loss = model.training_step(batch, 0)['loss']
    loss.backward()
    model.on_after_backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
names = sys.argv[1:] or ['aten::copy_', 'aten::add', 'aten::add_', 'aten::clone', 'aten::contiguous', 'aten::cat']
rows = [e for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6) if e.key in names and e.device_time_total > 50]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:40]:
    stack = [s for s in e.stack if 'sparse_vae_b200' in s or 'torch/autograd' in s][:3]
    print(f"{e.device_time_total / 1e3:8.3f} ms x{e.count:4d}  {e.key:18s} {str(e.input_shapes)[:90]}  {' | '.join(s.split('/')[-1][:60] for s in stack)}")
