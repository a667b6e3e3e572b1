"""Oracle (test infrastructure): functional CPU restatement of one `TransformerVAE.training_step`.

A plain-torch fp32 restatement of the reference's training step, written against a flat
`{parameter_name: tensor}` dict that uses the reference's own state_dict key names, so that weights
can be exchanged with the reference model (golden fixtures) and with the product model
(`sparse_vae_b200.TransformerVAE`, parity tests).  Each function cites what it follows:

  training_step        sparse_vae/transformer_vae.py:42-66
  reconstruct          sparse_vae/transformer_vae.py:85-93
  Perceiver            sparse_vae/core/perceiver.py:8-50
  TransformerLayer     sparse_vae/core/transformer_layer.py:44-61
  Attention            sparse_vae/core/attention.py:51-105 (rotary :194-208)
  SparseAttention      sparse_vae/core/sparse_attention.py:75-92 -> oracle.attention.dense_masked_attention
  ConditionalGaussian  sparse_vae/core/conditional_gaussian.py:18-30
  sample_z             sparse_vae/core/continuous_autoencoder.py:42-52
  get_nll              sparse_vae/core/language_model.py:98-113,161-170

It is also what `bench.py --impl reference` / `cpu_baseline` time on the host cores (kind "port").
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import attention as oattn
from . import layout as olayout


def rotary(x: torch.Tensor, start: int = 0, max_pos: int = 10000) -> torch.Tensor:
    """core/attention.py:194-208; positions/angles are computed in x.dtype like the reference."""
    d_half = x.shape[-1] // 2
    freqs = torch.arange(d_half, dtype=x.dtype, device=x.device)
    positions = torch.arange(start, start + x.shape[-2], dtype=x.dtype, device=x.device)
    theta = max_pos ** (-freqs / d_half)
    angles = positions[:, None] * theta
    xg = x.reshape(*x.shape[:-1], d_half, 2)
    x0, x1 = xg[..., 0], xg[..., 1]
    cos, sin = angles.cos(), angles.sin()
    out = torch.stack([x0 * cos - x1 * sin, x1 * cos + x0 * sin], dim=-1)
    return out.reshape(x.shape)


def _linear(p, prefix, x):
    return F.linear(x, p[prefix + '.weight'], p.get(prefix + '.bias'))


def _layer_norm(p, prefix, x):
    return F.layer_norm(x, x.shape[-1:], p[prefix + '.weight'], p[prefix + '.bias'])


def attention(p: Dict[str, torch.Tensor], prefix: str, q: Optional[torch.Tensor], k: torch.Tensor, v: torch.Tensor,
              num_heads: int, padding: Optional[torch.Tensor], causal: bool = False, sparse_window: int = 0,
              max_length: int = 10000, sparse_cfg: Optional[dict] = None) -> torch.Tensor:
    """`Attention.forward`.  `padding` is the [B, L_k] bool mask the reference reads off `k.padding`."""
    block = 32
    max_pos = max_length if not sparse_window else 2 * sparse_window * block
    lq_name = prefix + '.learned_queries'
    if lq_name in p:
        q = p[lq_name].expand(k.shape[0], *p[lq_name].shape[1:])
    else:
        q = rotary(_linear(p, prefix + '.q_linear', q), 0, max_pos)
    k, v = _linear(p, prefix + '.k_linear', k), _linear(p, prefix + '.v_linear', v)
    k = rotary(k, 0, max_pos)

    def split(t):
        return t.reshape(t.shape[0], t.shape[1], num_heads, -1).transpose(1, 2)

    q, k, v = split(q), split(k), split(v)
    if sparse_window:
        cfg = dict(causal=True, include_cls=True)
        cfg.update(sparse_cfg or {})
        nb = q.shape[-2] // block
        lay = olayout.layout_2d(nb, sparse_window, cfg['causal'], cfg['include_cls'])
        kpm = oattn.reference_kpm(padding) if padding is not None else None
        out = oattn.dense_masked_attention(q, k, v, lay, block, cfg['causal'], kpm)
    else:
        scores = q @ k.transpose(-1, -2) * k.shape[-1] ** -0.5
        mask = padding[:, None, None, :] if padding is not None else None
        if causal:
            ql = q.shape[-2]
            cm = torch.ones(ql, ql, dtype=torch.bool, device=q.device).triu(1)
            mask = (mask | cm) if mask is not None else cm
        if mask is not None:
            scores = scores - mask * 1e7
        out = scores.softmax(dim=-1) @ v
    out = out.transpose(1, 2).reshape(out.shape[0], out.shape[2], -1)
    return _linear(p, prefix + '.output_linear', out)


def transformer_layer(p, prefix, x, num_heads, padding, context=None, context_padding=None, causal=False,
                      sparse_window=0, dropout_p=0.0):
    """`TransformerLayer.forward`.  `padding` must already be None where the reference's
    `PaddedTensor.padding` would not fit the tensor (core/padded_tensor.py:54-69)."""
    y = _layer_norm(p, prefix + '.attn_layer_norm', x)
    y = attention(p, prefix + '.attention', y, y, y, num_heads, padding, causal, sparse_window)
    x = x + y if x.shape == y.shape else y
    if (prefix + '.cross_attention.k_linear.weight') in p and context is not None:
        c = _layer_norm(p, prefix + '.context_layer_norm', context)
        y = _layer_norm(p, prefix + '.cross_attn_layer_norm', x)
        y = attention(p, prefix + '.cross_attention', y, c, c, num_heads, context_padding)
        x = x + y
    y = _layer_norm(p, prefix + '.ffn_layer_norm', x)
    y = F.linear(F.gelu(_linear(p, prefix + '.ffn.0', y)), p[prefix + '.ffn.2.weight'])
    if dropout_p:
        y = F.dropout(y, dropout_p, training=True)
    return x + y


def _fits(padding, length):
    return padding if (padding is not None and padding.shape[-1] == length) else None


def perceiver(p, x, padding, d_model, dropout_p=0.0):
    """`Perceiver.forward` with bottleneck_width=1 (transformer_vae.py:34-36)."""
    heads = d_model // 64
    z = transformer_layer(p, 'encoder.first_layer', x, heads, padding, dropout_p=dropout_p)
    i = 0
    while f'encoder.middle_layers.{i}.attn_layer_norm.weight' in p:
        z = transformer_layer(p, f'encoder.middle_layers.{i}', z, heads, _fits(padding, z.shape[1]), context=x,
                              context_padding=padding, dropout_p=dropout_p)
        i += 1
    return transformer_layer(p, 'encoder.bottleneck', z, heads, _fits(padding, z.shape[1]), dropout_p=dropout_p)


def reconstruct(p, x, z, padding, num_heads, num_layers, window, sparse=True, dropout_p=0.0):
    for i in range(num_layers):
        z_hidden = _linear(p, f'z_projections.{i}', z)
        x = torch.cat([z_hidden, x[..., 1:, :]], dim=-2)
        x = transformer_layer(p, f'decoder_layers.{i}', x, num_heads, padding, causal=True,
                              sparse_window=window if sparse else 0, dropout_p=dropout_p)
    h = _linear(p, 'output_layer.0', x)
    h = _layer_norm(p, 'output_layer.2', F.gelu(h))
    return F.linear(h, p['input_layer.0.weight'], p['output_layer.3.bias'])


def training_step(p: Dict[str, torch.Tensor], token_ids: torch.Tensor, num_tokens: torch.Tensor, eps: torch.Tensor,
                  d_model: int, num_heads: int, num_layers: int, window: int = 4, kl_weight: float = 1.0,
                  sparse: bool = True, dropout_p: float = 0.0):
    """Returns dict(loss, nll, kl, raw_kl, z, mu, logvar, logits?).  `eps` is the rsample noise [B,1,latent]."""
    original = token_ids.long()
    padding = original.eq(0)
    x = F.embedding(original, p['input_layer.0.weight'])
    enc = perceiver(p, x, padding, d_model, dropout_p)
    mulogvar = _linear(p, 'q_of_z_given_x.linear', enc)
    mu, logvar = mulogvar.chunk(2, dim=-1)
    var = logvar.exp()
    sigma = var.sqrt()
    z = mu + eps * sigma
    kl_elem = 0.5 * (mu ** 2 + var - logvar - 1.0)
    raw_kl = kl_elem.flatten(1).sum(dim=-1)
    kl = (raw_kl / num_tokens).mean()
    logits = reconstruct(p, x, z, padding, num_heads, num_layers, window, sparse, dropout_p)[..., :-1, :]
    nll = F.cross_entropy(logits.flatten(end_dim=1), original[..., 1:].flatten(), ignore_index=0)
    loss = nll + kl_weight * kl
    return dict(loss=loss, nll=nll, kl=kl, raw_kl=raw_kl, z=z, mu=mu, logvar=logvar)


def init_params(d_model: int = 512, num_layers: int = 6, latent: int = 64, vocab: int = 2 ** 15,
                init_scale: float = 0.02, seed: int = 7295) -> Dict[str, torch.Tensor]:
    """Random-init parameter dict with the reference's names and shapes (transformer_vae.py:26-40,
    core/transformer_language_model.py:34-72, core/perceiver.py:8-28) and its BERT-style init
    (core/language_model.py:80-96): N(0, init_scale) weights, zero biases, LayerNorm at identity,
    learned queries ~ N(0, 1)."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}

    def lin(name, out_f, in_f, bias=True):
        p[name + '.weight'] = torch.randn(out_f, in_f, generator=g) * init_scale
        if bias:
            p[name + '.bias'] = torch.zeros(out_f)

    def ln(name):
        p[name + '.weight'] = torch.ones(d_model)
        p[name + '.bias'] = torch.zeros(d_model)

    def attn(name, learned=0):
        if learned:
            p[name + '.learned_queries'] = torch.randn(1, learned, d_model, generator=g)
        else:
            lin(name + '.q_linear', d_model, d_model)
        for nm in ('k_linear', 'v_linear', 'output_linear', 'pos_linear'):
            lin(f'{name}.{nm}', d_model, d_model)

    def layer(name, learned=0, cross=False):
        attn(name + '.attention', learned)
        lin(name + '.ffn.0', 4 * d_model, d_model)
        lin(name + '.ffn.2', d_model, 4 * d_model, bias=False)
        ln(name + '.attn_layer_norm')
        ln(name + '.ffn_layer_norm')
        if cross:
            attn(name + '.cross_attention')
            ln(name + '.cross_attn_layer_norm')
            ln(name + '.context_layer_norm')

    p['input_layer.0.weight'] = torch.randn(vocab, d_model, generator=g) * init_scale
    lin('output_layer.0', d_model, d_model)
    ln('output_layer.2')
    p['output_layer.3.bias'] = torch.zeros(vocab)
    for i in range(num_layers):
        layer(f'decoder_layers.{i}')
        lin(f'z_projections.{i}', d_model, latent)
    lin('q_of_z_given_x.linear', 2 * latent, d_model)
    enc_layers = num_layers // 2
    layer('encoder.first_layer', learned=64)
    layer('encoder.bottleneck', learned=1)
    for i in range(enc_layers - 2):
        layer(f'encoder.middle_layers.{i}', cross=True)
    return {k: v.requires_grad_(True) for k, v in p.items()}
