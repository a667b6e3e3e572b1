"""Wall / device time of model.sample(L, 256) (BASELINE config 5) -- run from the root of the tree to measure."""
import sys, time, os
from pathlib import Path
import torch
sys.path.insert(0, os.getcwd())
import sparse_vae_b200 as sv
from sparse_vae_b200.core.lightning_shim import to_attrdict
print('package', Path(sv.__file__).parent)
dev = torch.device('cuda')
torch.manual_seed(7295)
model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams())).to(dev).eval()
model.initialize_weights()
model.start_token, model.end_token = 1, 2
with torch.no_grad():
    model.sample(64, 256)
    for L in (1024, 4096, 4096):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ids = model.sample(L, 256)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f'L {L}: {dt * 1e3:8.1f} ms = {dt * 1e6 / (ids.shape[1] - 1):6.1f} us/token, ids {tuple(ids.shape)}')
