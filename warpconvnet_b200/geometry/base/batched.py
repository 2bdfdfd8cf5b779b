# SPDX-License-Identifier: Apache-2.0
"""Concatenated batch container: one ``[N, C]`` tensor + CPU ``offsets[B+1]``.

Mirrors the contract of warpconvnet/geometry/base/batched.py:14-270 (offsets always live on the
CPU, list-of-tensors or cat-tensor+offsets constructors).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor


def list_to_cat_tensor(tensors: Sequence[Tensor]) -> Tuple[Tensor, Tensor, None]:
    offsets = [0]
    for t in tensors:
        offsets.append(offsets[-1] + int(t.shape[0]))
    return torch.cat(list(tensors), dim=0), torch.LongTensor(offsets), None


class BatchedTensor:
    batched_tensor: Tensor
    offsets: Tensor

    def __init__(self, batched_tensor, offsets=None, device: Optional[str] = None):
        if isinstance(batched_tensor, (list, tuple)):
            assert offsets is None, "If batched_tensors is a list, offsets must be None."
            batched_tensor, offsets, _ = list_to_cat_tensor(batched_tensor)
        else:
            assert isinstance(batched_tensor, Tensor), "Batched tensor must be a tensor or a list"
            if offsets is None:
                offsets = [0, batched_tensor.shape[0]]
        if isinstance(offsets, (list, tuple)):
            offsets = torch.LongTensor(list(offsets))
        elif isinstance(offsets, Tensor):
            offsets = offsets.cpu()
        else:
            raise ValueError(f"Invalid offsets type {type(offsets)}")
        if device is not None:
            batched_tensor = batched_tensor.to(device)
        self.offsets = offsets
        self.batched_tensor = batched_tensor
        self.check()

    @property
    def batch_size(self) -> int:
        return len(self.offsets) - 1

    def check(self):
        assert self.offsets.device.type == "cpu" and self.offsets.dtype in (
            torch.int32, torch.int64), f"Offsets must be a cpu int tensor, got {self.offsets}"
        assert isinstance(self.batched_tensor, Tensor)

    def to(self, device=None, dtype=None):
        t = self.batched_tensor
        if device is not None:
            t = t.to(device)
        if dtype is not None:
            t = t.to(dtype)
        return self.__class__(t, self.offsets)

    @property
    def device(self):
        return self.batched_tensor.device

    @property
    def shape(self):
        return self.batched_tensor.shape

    @property
    def dtype(self):
        return self.batched_tensor.dtype

    def half(self):
        return self.to(dtype=torch.float16)

    def float(self):
        return self.to(dtype=torch.float32)

    def double(self):
        return self.to(dtype=torch.float64)

    def numel(self):
        return self.batched_tensor.numel()

    def __len__(self) -> int:
        return len(self.batched_tensor)  # rows, like the reference (batched.py:151-152)

    def __getitem__(self, idx: int) -> Tensor:
        return self.batched_tensor[int(self.offsets[idx]):int(self.offsets[idx + 1])]

    def equal_shape(self, value: "BatchedTensor") -> bool:
        return (len(self.offsets) == len(value.offsets)
                and bool((self.offsets == value.offsets).all()) and self.numel() == value.numel())

    def equal_rigorous(self, value: "BatchedTensor") -> bool:
        return (isinstance(value, BatchedTensor) and self.equal_shape(value)
                and bool((self.batched_tensor == value.batched_tensor).all()))

    def __eq__(self, value) -> bool:
        """Shape-level equality (offsets + element count), batched.py:178-182."""
        return isinstance(value, BatchedTensor) and self.equal_shape(value)

    __hash__ = object.__hash__

    def binary_op(self, value, op: str) -> "BatchedTensor":
        """Scalar / one-element tensor / same-shape BatchedTensor operand (batched.py:184-210)."""
        if isinstance(value, BatchedTensor):
            assert self.equal_shape(value), f"Shapes do not match: {self} vs {value}"
            value = value.batched_tensor
        elif not (isinstance(value, (int, float)) or (torch.is_tensor(value)
                                                      and value.numel() == 1)):
            raise TypeError(f"unsupported operand {type(value)} for {op}")
        return self.__class__(getattr(self.batched_tensor, op)(value), self.offsets)

    def to_nested(self) -> Tensor:
        return torch.nested.nested_tensor([self[i] for i in range(self.batch_size)],
                                          requires_grad=self.batched_tensor.requires_grad)

    @classmethod
    def from_nested(cls, nested: Tensor) -> "BatchedTensor":
        grad = nested.requires_grad
        return cls([t.requires_grad_(grad) for t in nested.unbind()])

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(offsets={self.offsets.tolist()}, "
                f"shape={tuple(self.shape)}, device={self.device}, dtype={self.dtype})")

    def __str__(self) -> str:
        return f"{self.__class__.__name__}(offsets={self.offsets.tolist()}, shape={tuple(self.shape)})"


for _name in ("add", "sub", "mul", "truediv", "floordiv", "mod", "pow"):
    setattr(BatchedTensor, f"__{_name}__",
            (lambda op: lambda self, value: self.binary_op(value, op))(f"__{_name}__"))


class Features(BatchedTensor):
    @property
    def num_channels(self):
        return self.batched_tensor.shape[-1]

    @property
    def is_cat(self):
        return True

    @property
    def is_pad(self):
        return False


class CatFeatures(Features):
    """warpconvnet/geometry/features/cat.py:11-34."""

    def check(self):
        super().check()
        assert self.batched_tensor.ndim == 2, "Batched tensor must be 2D"
        assert self.batched_tensor.shape[0] == int(self.offsets[-1]), (
            f"Offsets {self.offsets} does not match tensors {self.batched_tensor.shape}")

    def to_cat(self) -> "CatFeatures":
        return self

    def to_pad(self, pad_multiple: Optional[int] = None) -> "PadFeatures":
        """[B, Lmax, C] zero-padded copy (features/ops/convert.py cat_to_pad_tensor)."""
        counts = self.offsets.diff()
        width = int(counts.max()) if len(counts) else 0
        if pad_multiple:
            width = -(-width // pad_multiple) * pad_multiple
        out = self.batched_tensor.new_zeros((self.batch_size, width, self.num_channels))
        for b in range(self.batch_size):
            out[b, : int(counts[b])] = self[b]
        return PadFeatures(out, self.offsets, pad_multiple)


class PadFeatures(Features):
    """Zero-padded ``[B, Lmax, C]`` features (warpconvnet/geometry/features/pad.py); kept to
    convert to and from the concatenated layout — the sparse-conv path itself is cat-only."""

    def __init__(self, batched_tensor: Tensor, offsets, pad_multiple: Optional[int] = None,
                 device=None):
        self.pad_multiple = pad_multiple
        super().__init__(batched_tensor, offsets, device=device)

    def check(self):
        BatchedTensor.check(self)
        assert self.batched_tensor.ndim == 3, "Padded features must be [B, L, C]"
        assert self.batched_tensor.shape[0] == len(self.offsets) - 1

    @property
    def is_cat(self):
        return False

    @property
    def is_pad(self):
        return True

    @property
    def max_num_points(self) -> int:
        return self.batched_tensor.shape[1]

    def to(self, device=None, dtype=None):
        t = self.batched_tensor.to(device=device or self.device, dtype=dtype)
        return PadFeatures(t, self.offsets, self.pad_multiple)

    def __getitem__(self, idx: int) -> Tensor:
        return self.batched_tensor[idx, : int(self.offsets[idx + 1] - self.offsets[idx])]

    def to_pad(self, pad_multiple: Optional[int] = None) -> "PadFeatures":
        return self if pad_multiple == self.pad_multiple else self.to_cat().to_pad(pad_multiple)

    def to_cat(self) -> CatFeatures:
        return CatFeatures(torch.cat([self[b] for b in range(self.batch_size)], dim=0),
                           self.offsets)

    def equal_shape(self, value) -> bool:
        return (isinstance(value, PadFeatures) and self.shape == value.shape
                and bool((self.offsets == value.offsets).all()))


def to_batched_features(features, offsets, device=None) -> CatFeatures:
    if isinstance(features, Tensor):
        if features.ndim != 2:
            raise ValueError(f"Invalid features tensor shape {features.shape}")
        return CatFeatures(features, offsets, device=device)
    if isinstance(features, Features):
        return features.to(device) if device is not None else features
    raise TypeError(f"Features must be a tensor or CatFeatures, got {type(features)}")


class Coords(BatchedTensor):
    @property
    def num_spatial_dims(self):
        return self.batched_tensor.shape[1]
