"""The slice of `omegaconf` the reference's train.py uses (train.py:17-24,55-61), on this package's attribute dict:
`OmegaConf.create(dict)`, `OmegaConf.structured(dataclass)`, `cfg.merge_with_dotlist(["a.b=1", ...])`,
`cfg.merge_with(other)`, `cfg.get(key)`, attribute access and `**cfg` expansion."""
from __future__ import annotations

import dataclasses
from typing import Any

from ..core.lightning_shim import AttrDict, to_attrdict


def _parse_value(text: str) -> Any:
    low = text.strip().lower()
    if low in ('true', 'false'):
        return low == 'true'
    if low in ('null', 'none', '~'):
        return None
    for cast in (int, float):
        try:
            return cast(text)
        except ValueError:
            pass
    if text.startswith('[') and text.endswith(']'):
        return [_parse_value(t) for t in text[1:-1].split(',') if t.strip()]
    return text


class DictConfig(AttrDict):
    def __getattr__(self, key):            # OmegaConf returns None for a missing key of an untyped node
        try:
            return self[key]
        except KeyError:
            if key.startswith('__'):
                raise AttributeError(key) from None
            return None

    def merge_with(self, *others):
        for other in others:
            for k, v in dict(other).items():
                if isinstance(v, dict) and isinstance(self.get(k), dict):
                    DictConfig.merge_with(self[k], v) if isinstance(self[k], DictConfig) else self[k].update(v)
                else:
                    self[k] = _wrap(v)

    def merge_with_dotlist(self, dotlist):
        for item in dotlist:
            key, _, value = item.partition('=')
            node = self
            *parents, leaf = key.split('.')
            for part in parents:
                if not isinstance(node.get(part), dict):
                    node[part] = DictConfig()
                node = node[part]
            node[leaf] = _parse_value(value)


def _wrap(value):
    if isinstance(value, dict) and not isinstance(value, DictConfig):
        return DictConfig({k: _wrap(v) for k, v in value.items()})
    return value


class OmegaConf:
    @staticmethod
    def create(obj=None) -> DictConfig:
        return _wrap(dict(obj or {}))

    @staticmethod
    def structured(obj) -> DictConfig:
        if dataclasses.is_dataclass(obj):
            return _wrap(dict(to_attrdict(obj)))
        return _wrap(dict(obj))

    @staticmethod
    def to_container(cfg, resolve: bool = True):
        return {k: OmegaConf.to_container(v) if isinstance(v, dict) else v for k, v in dict(cfg).items()}
