"""Is the training step CPU-bound?  Host time to ENQUEUE a step versus device time to run it (debug; not a bench)."""
import sys
import time
from pathlib import Path

import torch

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
model.validate_posterior = False
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
# (a) host enqueue time with an empty GPU queue in front (sync before each step, time until step() returns)
enq = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    enq.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
# (b) steady state, no syncs
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3 / 5
print(f'host time to enqueue one step: {sorted(enq)[len(enq) // 2]:.1f} ms (median of {[round(x, 1) for x in enq]})')
print(f'steady-state wall time per step: {wall:.1f} ms')
import cProfile, pstats
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
