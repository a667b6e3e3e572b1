#!/usr/bin/env python
"""Encodes a synthetic batch to its posterior mean and decodes it back teacher-forced through the block-sparse
decoder (the inference-mode user of the sparse kernel: TransformerVAE.predict + reconstruct)."""
import torch

import sparse_vae_b200 as sv
from sparse_vae_b200.core.lightning_shim import to_attrdict
from sparse_vae_b200.synthetic import synthetic_tokens, to_device

if __name__ == '__main__':
    model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams())).cuda().eval()
    model.initialize_weights()
    batch = to_device(synthetic_tokens(2, 1024), torch.device('cuda'))
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
        posterior = model.predict(batch)
        x = model.input_layer(batch['token_ids'].as_raw().long())
        logits = model.reconstruct(x, posterior.loc, padding=batch['token_ids'].padding)
    print(logits.argmax(-1)[:, :16])
