"""Which Python call sites issue the big aten::copy_ / add_ / mul ops of one eager training step?"""
import sys
from pathlib import Path
from collections import defaultdict
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import sparse_vae_b200 as sv
from sparse_vae_b200.core.lightning_shim import to_attrdict
from sparse_vae_b200.data_parallel import GradientAllReducer
from sparse_vae_b200.synthetic import synthetic_tokens, to_device
dev = torch.device('cuda'); B, L = 16, 4096
torch.manual_seed(7295)
model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams())).to(dev); model.initialize_weights()
(opt,), _ = model.configure_optimizers(tokens_per_batch=B * L)
reducer = GradientAllReducer(model)
batch = to_device(synthetic_tokens(B, L), dev)
def step():
    reducer.zero_grad()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        loss = model.training_step(batch, 0)['loss']
    loss.backward(); reducer.finish(); model.on_after_backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step(); torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.name in ('aten::copy_', 'aten::add_', 'aten::mul', 'aten::add', 'aten::fill_', 'aten::gather') and e.self_device_time_total > 15:
        site = next((s for s in e.stack if 'sparse_vae_b200' in s or 'bench' in s), (e.stack[0] if e.stack else '?'))
        agg[(e.name, str(e.input_shapes)[:60], site[-95:])][0] += 1
        agg[(e.name, str(e.input_shapes)[:60], site[-95:])][1] += e.self_device_time_total
for (name, shapes, site), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f'{us / 1e3:7.3f} ms {n:3d} x {name:12s} {shapes:60s} {site}')
