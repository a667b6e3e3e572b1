"""One `sample()` of 256 samples x N tokens with the default hparams (for ncu captures of the decoding kernels)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sparse_vae_b200 as sv  # noqa: E402
from sparse_vae_b200.core.lightning_shim import to_attrdict  # noqa: E402

torch.manual_seed(7295)
model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams())).cuda().eval()
model.initialize_weights()
model.start_token, model.end_token = 1, 2
with torch.no_grad():
    ids = model.sample(int(sys.argv[1]) if len(sys.argv) > 1 else 240, 256)
torch.cuda.synchronize()
print(ids.shape)
