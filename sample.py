#!/usr/bin/env python
"""`python sample.py [max_length] [batch_size]`: draws z ~ N(0, I) and decodes autoregressively with the KV-cached
decoder, like the reference's sample.py -> TransformerVAE.sample (random-init weights unless a state_dict path is
given as third argument; reference checkpoints load because the parameter names are identical)."""
import sys

import torch

import sparse_vae_b200 as sv
from sparse_vae_b200.core.lightning_shim import to_attrdict

if __name__ == '__main__':
    max_length = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    batch_size = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams())).cuda().eval()
    model.initialize_weights()
    if len(sys.argv) > 3:
        state = torch.load(sys.argv[3], map_location='cuda')
        model.load_state_dict(state.get('state_dict', state))
    model.start_token, model.end_token = 1, 2
    with torch.no_grad():
        print(model.sample(max_length, batch_size))
