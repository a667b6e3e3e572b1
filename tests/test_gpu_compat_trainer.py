"""GPU: the path the reference's train.py takes (train.py:17-24,78,94-95) -- OmegaConf-style config, `TextDataModule`,
`Trainer(accumulate_grad_batches=2, precision=16).fit(model, datamodule=data)` -- with this package's TransformerVAE
on the sparse kernels.  (The script itself cannot travel to the GPU box; tests/test_compat_reference_scripts.py runs it,
unchanged, in the build container.)"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('precision', [16, 'bf16'])
def test_trainer_fit_transformer_vae(precision):
    from sparse_vae_b200 import compat
    compat.install()
    from omegaconf import OmegaConf
    from pytorch_lightning import Trainer, seed_everything
    from sparse_vae import TextDataModule, TransformerVAE, TransformerVAEHparams

    seed_everything(7295)
    config = OmegaConf.create({'trainer': {'accumulate_grad_batches': 2, 'checkpoint_callback': False, 'precision': precision}})
    config.model = OmegaConf.structured(TransformerVAEHparams)
    config.merge_with_dotlist(['model.d_model=256', 'model.num_layers=4', 'model.latent_depth=32', 'trainer.max_steps=4',
                               'trainer.gpus=[0]', 'data.tokens_per_batch=2048', 'data.seq_len=512', 'trainer.log_every_n_steps=1'])
    model = TransformerVAE(config.model)
    data = TextDataModule(**config.get('data', {}))
    trainer = Trainer(**config.trainer, logger=False)
    trainer.fit(model, datamodule=data)
    assert trainer.global_step == 4 and model.global_step == 4
    m = trainer.logged_metrics
    assert all(math.isfinite(m[k]) for k in ('loss', 'train_nll', 'train_kl', 'grad_norm')), m
    assert 9.0 < m['train_nll'] < 11.5, m                       # ln(32768) = 10.4 at initialisation
    assert next(model.parameters()).is_cuda
    for p in model.parameters():
        assert torch.isfinite(p).all()
