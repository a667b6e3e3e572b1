"""Minimal stand-ins for the two third-party packages the reference's model classes lean on
(pytorch_lightning.LightningModule and omegaconf.DictConfig), neither of which is installed in this image.
Only what the hot path's callers touch is provided: hyper-parameter storage, `self.log`, `self.device`,
`global_step`, and no-op hooks.  If the real packages are importable they are used instead.
"""
from __future__ import annotations

import dataclasses
from typing import Any

import torch
from torch import nn

try:                                                    # pragma: no cover - not available in this image
    import pytorch_lightning as pl                      # type: ignore
    LightningModule = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:                                       # noqa: BLE001
    HAVE_LIGHTNING = False

    class LightningModule(nn.Module):
        def __init__(self):
            super().__init__()
            self._hparams = AttrDict()
            self.logged = {}
            self.global_step = 0
            self.trainer = None

        @property
        def hparams(self):
            return self._hparams

        def save_hyperparameters(self, hparams=None):
            if hparams is not None:
                self._hparams = to_attrdict(hparams)

        def log(self, name: str, value: Any, **kwargs):
            self.logged[name] = value.detach() if isinstance(value, torch.Tensor) else value

        @property
        def device(self) -> torch.device:
            return next(self.parameters()).device

        # hooks the training loop calls
        def on_fit_start(self): ...
        def on_train_start(self): ...
        def on_after_backward(self): ...


class AttrDict(dict):
    """dict with attribute access (the part of omegaconf.DictConfig the models use)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        self[key] = value


try:                                                    # pragma: no cover
    from omegaconf import DictConfig                    # type: ignore
except Exception:                                       # noqa: BLE001
    DictConfig = AttrDict


def to_attrdict(hparams) -> AttrDict:
    if isinstance(hparams, AttrDict):
        return hparams
    if dataclasses.is_dataclass(hparams) and not isinstance(hparams, type):
        return AttrDict(dataclasses.asdict(hparams))
    if dataclasses.is_dataclass(hparams):
        return AttrDict(dataclasses.asdict(hparams()))
    return AttrDict(dict(hparams))
