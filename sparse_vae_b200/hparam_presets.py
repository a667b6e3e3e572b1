"""Named hyper-parameter presets, `{'data': ..., 'model': ..., 'trainer': ...}` per preset: the same eight
names and values as the reference's hparam_presets.py, assembled from shared pieces instead of spelled out.
tests/test_host_surface.py checks the assembled dict against a fixture dumped from the reference."""
from __future__ import annotations


def _wiki(tokens, max_tokens):
    return dict(dataset_name='wikipedia', dataset_config='20200501.en', tokens_per_batch=tokens,
                min_tokens_per_sample=512, max_tokens_per_sample=max_tokens)


def _pg19(tokens, max_tokens):
    return dict(dataset_name='pg19', dataset_config=None, tokens_per_batch=tokens,
                min_tokens_per_sample=512, max_tokens_per_sample=max_tokens)


def _transformer(sparse, **extra):
    base = dict(d_model=512, grad_checkpointing=True, grad_clip_threshold=150.0, init_scale=0.02, lr=3e-4,
                num_layers=6, sparse_self_attention=sparse, tie_embedding_weights=True)
    base.update(extra)
    return base


def _kl(start, end=None, steps=8000):
    out = dict(kl_weight_start=start, kl_annealing_steps=steps, latent_depth=64)
    if end is not None:
        out['kl_weight_end'] = end
    return out


def _lstm(d_model, kl_start, kl_steps):
    return dict(bidirectional_encoder=True, d_model=d_model, d_embedding=512, grad_clip_threshold=150.0,
                init_scale=None, kl_weight_start=kl_start, kl_annealing_steps=kl_steps, latent_depth=64, lr=3e-4,
                tie_embedding_weights=True, tie_logit_weights=True, transformer_encoder=False)


def _trainer(accum, val_check=None):
    out = dict(accumulate_grad_batches=accum)
    if val_check is not None:
        out['val_check_interval'] = val_check
    return out


hparam_presets = {
    'lstm-benchmark': {'model': _lstm(1024, 0.2, 8000), 'trainer': _trainer(2)},
    'lstm-wikipedia': {'data': _wiki(50_000, 25_000), 'model': _lstm(2048, 1.0, 0), 'trainer': _trainer(2, 0.25)},
    'dense-benchmark': {'data': _wiki(50_000, 3_125), 'model': _transformer(False, **_kl(0.3, 1.0)),
                        'trainer': _trainer(2)},
    'sparse-benchmark': {'data': _wiki(50_000, 3_125), 'model': _transformer(True, **_kl(1.0, steps=0)),
                         'trainer': _trainer(2)},
    'nonvae-wikipedia': {'data': _wiki(50_000, 3_125), 'model': _transformer(False), 'trainer': _trainer(2, 0.1)},
    'wikipedia': {'data': _wiki(100_000, 50_000), 'model': _transformer(True, attn_window_size=8, **_kl(0.1, 1.0)),
                  'trainer': _trainer(2, 0.1)},
    'pg19': {'data': _pg19(102_912, 102_400), 'model': _transformer(True, attn_window_size=6, **_kl(0.1, 1.0)),
             'trainer': _trainer(4, 0.5)},
    'nonvae-pg19': {'data': _pg19(92_672, 92_160), 'model': _transformer(True), 'trainer': _trainer(4, 0.5)},
}
