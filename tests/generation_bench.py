"""Times BASELINE config 5 on one GPU (not a bench.py line): (i) 256 x 4096 tokens through one teacher-forced decoder
forward per chunk of 32 samples, latents from the prior; (ii) `sample()` -- the reference's KV-cached autoregressive
decoding -- for a bounded number of tokens.  Prints one JSON line; run `python tests/generation_bench.py`."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sparse_vae_b200 as sv  # noqa: E402
from sparse_vae_b200.core.lightning_shim import to_attrdict  # noqa: E402
from sparse_vae_b200.synthetic import synthetic_tokens, to_device  # noqa: E402


def main():
    dev = torch.device('cuda')
    torch.manual_seed(7295)
    model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams())).to(dev).eval()
    model.initialize_weights()
    model.start_token, model.end_token = 1, 2
    chunks = [to_device(synthetic_tokens(32, 4096, seed=chunk), dev)['token_ids'] for chunk in range(8)]

    def decode_all():
        z_all = torch.randn(256, 1, model.hparams.latent_depth, device=dev)
        out = []
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
            for c, tokens in enumerate(chunks):
                x = model.input_layer(tokens.as_raw().long())
                logits = model.reconstruct(x, z_all[32 * c:32 * c + 32], padding=tokens.padding)
                out.append(logits.argmax(-1))
                del logits
        return torch.cat(out)

    decode_all()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    ids = decode_all()
    stop.record()
    torch.cuda.synchronize()
    fwd_ms = start.elapsed_time(stop)

    ar_tokens, ar_batch = 1024, 256
    with torch.no_grad():
        model.sample(8, ar_batch)
        torch.cuda.synchronize()
        start.record()
        sampled = model.sample(ar_tokens, ar_batch)
        stop.record()
        torch.cuda.synchronize()
    ar_ms = start.elapsed_time(stop)
    print(json.dumps({'config': 'C5: 256 samples x 4096 tokens, default hparams, bf16, 1 GPU',
                      'decoder_forward_ms': round(fwd_ms, 2),
                      'decoder_forward_tokens_per_s': round(ids.numel() / fwd_ms * 1e3),
                      'autoregressive_sample': {'batch': ar_batch, 'steps': int(sampled.shape[1]),
                                                'ms': round(ar_ms, 2),
                                                'tokens_per_s': round(sampled.numel() / ar_ms * 1e3)}}))


if __name__ == '__main__':
    main()
