"""`PaddedTensor`: a Tensor that carries the [batch, length] padding mask of its batch
(reference surface: sparse_vae/core/padded_tensor.py).  Re-implemented on the classmethod
`__torch_function__` protocol; the mask rides along through torch functions and `.padding` only hands it
out while it still fits the tensor's leading dimensions.

The model code in this package extracts the mask once and passes it down explicitly (plain tensors are
much cheaper to dispatch than a Python-level subclass on every op); `PaddedTensor` remains the batch-dict
interchange type of the reference's data module (text_data_module.py:194-210).
"""
from __future__ import annotations

from typing import Optional

from torch import Tensor


class PaddedTensor(Tensor):
    _padding: Optional[Tensor] = None

    @classmethod
    def from_raw(cls, data: Tensor, padding: Optional[Tensor] = None) -> 'PaddedTensor':
        out = data.as_subclass(cls)
        out.padding = data.as_subclass(Tensor).eq(0) if padding is None else padding
        return out

    @classmethod
    def unpadded(cls, data: Tensor) -> 'PaddedTensor':
        out = data.as_subclass(cls)
        out._padding = None
        return out

    def as_raw(self) -> Tensor:
        return self.as_subclass(Tensor)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        result = super().__torch_function__(func, types, args, kwargs)
        mask = None
        for a in _flatten(args):
            if isinstance(a, PaddedTensor) and a.__dict__.get('_padding') is not None:
                mask = a.__dict__['_padding']
                break
        if mask is not None:
            for r in _flatten((result,)):
                if isinstance(r, PaddedTensor) and r.__dict__.get('_padding') is None:
                    r.__dict__['_padding'] = mask
        return result

    @property
    def padding(self) -> Optional[Tensor]:
        mask = self.__dict__.get('_padding')
        if mask is None:
            return None
        if mask.device != self.device:
            mask = self.__dict__['_padding'] = mask.to(self.device)
        if mask.shape[0] == 1 and self.shape[0] != 1:
            return mask.expand(self.shape[0], *mask.shape[1:])
        # a mask that no longer matches the sequence dimension is not handed out
        if mask.ndim <= self.ndim and mask.shape[-1] == self.shape[mask.ndim - 1]:
            return mask
        return None

    @padding.setter
    def padding(self, value: Optional[Tensor]):
        if value is not None:
            assert value.ndim <= self.ndim, "Padding cannot have more dimensions than the tensor itself"
            for dim, (have, want) in enumerate(zip(value.shape, self.shape)):
                assert have == want, f"Padding size {have} must match data size {want} at dim {dim}"
            value = value.as_subclass(Tensor).to(self.device)
        self.__dict__['_padding'] = value

    def __repr__(self):
        mask = self.__dict__.get('_padding')
        return f"PaddedTensor(shape={list(self.shape)}, padding={None if mask is None else list(mask.shape)})"


def _flatten(xs):
    for x in xs:
        if isinstance(x, (list, tuple)):
            yield from _flatten(x)
        else:
            yield x


def split_padding(x):
    """(plain tensor, padding mask or None) of a tensor that may be a PaddedTensor."""
    if isinstance(x, PaddedTensor):
        return x.as_raw(), x.padding
    return x, None
