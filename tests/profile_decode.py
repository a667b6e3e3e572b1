"""Kernel-time breakdown of graphed token-by-token decoding (256 samples, default hparams): python tests/profile_decode.py"""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sparse_vae_b200 as sv  # noqa: E402
from sparse_vae_b200.core.lightning_shim import to_attrdict  # noqa: E402

dev = torch.device('cuda')
torch.manual_seed(7295)
model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams())).to(dev).eval()
model.initialize_weights()
model.start_token, model.end_token = 1, 2
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
with torch.no_grad():
    model.sample(16, 256)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model.sample(steps, 256)
        torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
print(f"total device time {total / 1e3:.1f} ms over {steps - 2} graphed steps = {total / (steps - 2):.1f} us/token")
for e in rows[:28]:
    print(f"{e.device_time_total / (steps - 2):9.1f} us/token  x{e.count / (steps - 2):6.1f}  {e.key[:110]}")

# ---- where the wall time goes: back-to-back replays vs the per-token host check ------------------------------------
import time  # noqa: E402

from sparse_vae_b200.core import decode  # noqa: E402
from sparse_vae_b200.core.attention import Attention  # noqa: E402
from sparse_vae_b200.core.generation import GenerationState  # noqa: E402

with torch.no_grad(), torch.autocast('cuda'):
    state = GenerationState(4096, 256, 1, 2, device=dev)
    z = torch.randn(256, 1, model.hparams.latent_depth, device=dev)
    with Attention.kv_cache(4096):
        x = model.input_layer(state.prev_tokens())
        for layer, project in zip(model.decoder_layers, model.z_projections):
            x = layer(project(z).to(x.dtype))
        Attention.update_kv_cache(state.process_logits(model.output_layer(x.squeeze(1))))
        dec = decode.GraphedDecoder(model, state)
        dec._capture()
        for _ in range(200):
            dec.graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(200):
            dec.graph.replay()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"back-to-back: device {e0.elapsed_time(e1) * 5:.1f} us/replay, host enqueue {(t1 - t0) * 5e3:.1f} us/replay, "
              f"wall {(t2 - t0) * 5e3:.1f} us/replay")
        stream = torch.cuda.current_stream()
        t0 = time.perf_counter()
        for _ in range(200):
            dec.graph.replay()
            dec.finished_host.copy_(dec.finished, non_blocking=True)
            stream.synchronize()
            done = int(dec.finished_host[0])
        t1 = time.perf_counter()
        print(f"replay + flag copy + sync per token: wall {(t1 - t0) * 5e3:.1f} us/token")
