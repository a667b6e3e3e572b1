"""CPU: host-side logic and the C-ABI boundary (no compute calls -- there is no GPU here)."""
import ctypes
import dataclasses
import json
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import sparse_vae_b200 as sv
from sparse_vae_b200 import _native as N
from sparse_vae_b200.core.lightning_shim import to_attrdict
from oracle import layout as olayout

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / 'include' / 'sparse_vae_b200.h').read_text()
    declared = set(re.findall(r'SVAE_API[^;]*?\b(svae_\w+)\s*\(', header))
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    lib = ctypes.CDLL(str(N.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert N.lib.svae_abi_version() == 1
    assert ctypes.sizeof(N.AttnDesc) == 12 * 4 + 8 * 3 * 8


def test_layout_from_library_is_bit_exact(golden_dir):
    g = np.load(golden_dir / 'layout_golden.npz')
    for i, (nb, w, causal, cls, nnz) in enumerate(g['meta']):
        ref = np.unpackbits(g[f'l{i}'])[:nb * nb].reshape(nb, nb).astype(np.int64)
        sa = sv.SparseAttention(window_size=int(w), causal=bool(causal), include_cls=bool(cls), num_heads=3)
        lay = sa.get_layout(int(nb))
        assert lay.dtype == torch.int64 and tuple(lay.shape) == (3, nb, nb)
        for h in range(3):
            assert np.array_equal(lay[h].numpy(), ref)
        assert sa.num_nonzero_blocks(int(nb)) == nnz
        row_ptr, col_idx, colT_ptr, rowT_idx = sa.get_lut(int(nb))
        rp, ci = olayout.csr(ref)
        cp, ri = olayout.csc(ref)
        assert np.array_equal(row_ptr.numpy(), rp) and np.array_equal(col_idx.numpy(), ci)
        assert np.array_equal(colT_ptr.numpy(), cp) and np.array_equal(rowT_idx.numpy(), ri)


def test_master_layout_shape_and_slice():
    sa = sv.SparseAttention(max_seq_len=32 * 40, num_heads=2)
    m = sa.get_master_layout()
    assert tuple(m.shape) == (2, 40, 40) and m.dtype == torch.int64
    assert torch.equal(m[..., :7, :7], sa.get_layout(7))
    assert sa.get_master_layout() is m                      # lru-cached like the reference


def test_sparse_attention_dataclass_contract():
    sa = sv.SparseAttention()
    assert dataclasses.asdict(sa) == dict(block_size=32, causal=True, include_cls=True, num_heads=8,
                                          max_seq_len=115_200, window_size=4)
    assert hash(sa) == hash(sv.SparseAttention()) and sa == sv.SparseAttention()
    with pytest.raises(dataclasses.FrozenInstanceError):
        sa.window_size = 5
    with pytest.raises(AssertionError):
        sv.SparseAttention(max_seq_len=100)
    q = torch.zeros(1, 8, 64, 64)
    with pytest.raises(ValueError, match='Only GPU devices are supported'):
        sa(q, q, q)                                         # no CPU path, like the reference's validator
    with pytest.raises(AssertionError):
        sa(q, q[..., :32, :], q)


def test_bottleneck_has_no_cpu_path():
    cg = sv.ConditionalGaussian(16, 4)
    with pytest.raises(ValueError, match='Only GPU devices are supported'):
        cg(torch.zeros(2, 1, 16))


def test_presets_match_reference(golden_dir):
    want = json.loads((golden_dir / 'presets_golden.json').read_text())
    assert json.loads(json.dumps(sv.hparam_presets)) == want


def test_hparam_defaults():
    hp = sv.TransformerVAEHparams()
    assert (hp.d_model, hp.num_heads, hp.num_layers, hp.attn_window_size, hp.sparse_self_attention) == (512, 8, 6, 4, True)
    assert (hp.latent_depth, hp.kl_weight, hp.lr, hp.init_scale, hp.grad_clip_threshold) == (64, 1.0, 2e-4, 0.02, 5.0)
    assert hp.tie_embedding_weights and not hp.grad_checkpointing and hp.early_stopping_metric == 'val_nll'


def test_state_dict_names_match_reference(golden_dir):
    g = np.load(golden_dir / 'model_golden.npz')
    want = {n: tuple(int(x) for x in s.split(',')) for n, s in zip(g['param_names'], g['param_shapes'])}
    model = sv.TransformerVAE(to_attrdict(sv.TransformerVAEHparams(d_model=256, num_layers=4)))
    have = {n: tuple(p.shape) for n, p in model.named_parameters()}
    assert have == want
    # tied embedding: one tensor behind three state_dict keys
    sd = model.state_dict()
    assert sd['input_layer.0.weight'].data_ptr() == sd['output_layer.3.weight'].data_ptr() == \
        sd['encoder_input_layer.0.weight'].data_ptr()


def test_padded_tensor_semantics():
    tok = torch.tensor([[5, 6, 0, 0], [7, 8, 9, 0]])
    pt = sv.PaddedTensor.from_raw(tok)
    assert torch.equal(pt.padding, tok.eq(0))
    emb = torch.nn.Embedding(10, 6)
    x = emb(pt)
    assert isinstance(x, sv.PaddedTensor) and torch.equal(x.padding, tok.eq(0))
    y = torch.nn.functional.layer_norm(x + 1, (6,))
    assert torch.equal(y.padding, tok.eq(0))
    assert y[:, :2].padding is None                        # mask no longer fits the sequence dimension
    assert type(pt.as_raw()) is torch.Tensor


def test_rotary_matches_oracle_restatement():
    from oracle.model import rotary
    x = torch.randn(2, 37, 16)
    for max_pos in (256, 10000):
        assert torch.allclose(sv.encode_position_rotary(x.clone(), 3, max_pos), rotary(x, 3, max_pos), atol=1e-6)
    xb = x.bfloat16()
    assert torch.equal(sv.encode_position_rotary(xb, 0, 256), rotary(xb, 0, 256))


def test_radam_matches_per_tensor_restatement():
    # restatement of the reference update rule (core/rectified_adam.py:15-88) for the non-LAMB branch
    torch.manual_seed(0)
    p = [torch.randn(5, 3), torch.randn(7)]
    params = [torch.nn.Parameter(t.clone()) for t in p]
    opt = sv.RAdam(params, lr=1e-2, weight_decay=0.01)
    ref = [t.clone().double() for t in p]
    m = [torch.zeros_like(t) for t in ref]
    v = [torch.zeros_like(t) for t in ref]
    b1, b2, eps, wd = 0.9, 0.999, 1e-6, 0.01
    for step in range(1, 9):
        grads = [torch.randn_like(t) for t in p]
        for prm, g in zip(params, grads):
            prm.grad = g.clone()
        opt.step()
        b2t = b2 ** step
        bias_v = (1 - b2t) ** 0.5
        rho_inf = 2 / (1 - b2) - 1
        rho_t = rho_inf - 2 * step * b2t / (1 - b2t)
        lr = 1e-2
        if rho_t > 4:
            lr *= (((rho_t - 4) * (rho_t - 2) * rho_inf) / ((rho_inf - 4) * (rho_inf - 2) * rho_t)) ** 0.5 * bias_v
        for i, g in enumerate(grads):
            g = g.double()
            m[i] = m[i] * b1 + (1 - b1) * g
            v[i] = v[i] * b2 + (1 - b2) * g * g
            ref[i] = ref[i] * (1 - lr * wd)
            step_size = lr / (1 - b1 ** step)
            if rho_t > 4:
                ref[i] = ref[i] - step_size * m[i] / (v[i].sqrt() / bias_v + eps)
            else:
                ref[i] = ref[i] - step_size * m[i]
    for prm, r in zip(params, ref):
        assert torch.allclose(prm.detach().double(), r, rtol=1e-5, atol=1e-6)
