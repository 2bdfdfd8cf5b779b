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
        return self.batch_size

    def __getitem__(self, idx: int) -> Tensor:
        return self.batched_tensor[int(self.offsets[idx]):int(self.offsets[idx + 1])]

    def equal_shape(self, value: "BatchedTensor") -> bool:
        return bool((self.offsets == value.offsets).all()) and self.numel() == value.numel()

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(offsets={self.offsets.tolist()}, shape={tuple(self.shape)})"


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
