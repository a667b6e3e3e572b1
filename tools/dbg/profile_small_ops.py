"""Device time of the non-GEMM ATen ops of one eager training step, grouped by op and operand shapes."""
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
skip = ('aten::mm', 'aten::addmm', 'aten::addmm_', 'aten::linear', 'aten::matmul')
for e in prof.events():
    if e.name.startswith('aten::') and e.name not in skip and e.self_device_time_total > 0:
        agg[(e.name, str(e.input_shapes)[:90])][0] += 1
        agg[(e.name, str(e.input_shapes)[:90])][1] += e.self_device_time_total
tot = sum(v[1] for v in agg.values())
print(f'non-GEMM ATen ops: {tot / 1e3:.3f} ms of device time, {sum(v[0] for v in agg.values())} calls')
for (name, shapes), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f'{us / 1e3:7.3f} ms  {n:4d} x {us / n:8.1f} us  {name:28s} {shapes}')
