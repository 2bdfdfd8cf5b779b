# SPDX-License-Identifier: Apache-2.0
"""Geometry base container: coordinates + features + extra attributes (incl. ``_cache``).

Same surface as warpconvnet/geometry/base/geometry.py:37-388 for the members the sparse-conv
path uses: ``replace`` (carries ``_extra_attributes`` — and therefore the kernel-map cache —
forward), ``feature_tensor`` (AMP aware), ``coordinate_tensor``, ``offsets``, ``cache``.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
from torch import Tensor

from .batched import CatFeatures, Coords, Features, to_batched_features


# caches whose entries index the rows of one particular coordinate set (kernel maps, strided
# coordinates, spatial grids): dropped whenever a new Geometry is built on other rows / devices
_ROW_CACHES = ("_cache", "_stride_cache", "_spatial_cache")


class Geometry:
    def __init__(self, batched_coordinates, batched_features, **kwargs):
        offsets = kwargs.pop("offsets", None)
        if isinstance(batched_coordinates, Tensor):
            assert offsets is not None, "offsets must be provided when coordinates is a tensor"
            batched_coordinates = Coords(batched_coordinates, offsets)
        self.batched_coordinates = batched_coordinates
        self.batched_features = to_batched_features(batched_features, batched_coordinates.offsets,
                                                    device=kwargs.get("device", None))
        kwargs.pop("device", None)
        assert bool((batched_coordinates.offsets == self.batched_features.offsets).all())
        if "_extra_attributes" in kwargs:
            attr = kwargs.pop("_extra_attributes")
            assert isinstance(attr, dict)
            kwargs.update(attr)
        self._extra_attributes: Dict[str, Any] = kwargs

    def __getitem__(self, idx: int) -> "Geometry":
        coords = self.batched_coordinates[idx]
        feats = self.batched_features[idx]
        attrs = {k: v for k, v in self._extra_attributes.items() if k not in _ROW_CACHES}
        return self.__class__(coords, feats, offsets=torch.tensor([0, len(coords)]), **attrs)

    def to(self, device=None, dtype=None) -> "Geometry":
        coords = self.batched_coordinates.to(device=device)
        feats = self.batched_features.to(device=device, dtype=dtype)
        attrs = {k: v for k, v in self._extra_attributes.items() if k not in _ROW_CACHES}
        return self.__class__(coords, feats, **attrs)

    @property
    def num_spatial_dims(self) -> int:
        return self.batched_coordinates.num_spatial_dims

    @property
    def coordinate_tensor(self) -> Tensor:
        return self.batched_coordinates.batched_tensor

    coordinates = coordinate_tensor

    @property
    def batch_indexed_coordinates(self) -> Tensor:
        from warpconvnet_b200.geometry.coords.ops.batch_index import batch_indexed_coordinates
        return batch_indexed_coordinates(self.coordinate_tensor, self.offsets)

    @property
    def coords(self) -> Tensor:
        """``[N, 1 + D]`` with the batch index in column 0 (geometry.py ``coords`` alias)."""
        return self.batch_indexed_coordinates

    @property
    def nested_coordinates(self) -> Tensor:
        return self.batched_coordinates.to_nested()

    @property
    def feature_tensor(self) -> Tensor:
        """Under autocast the features are returned in the autocast dtype
        (geometry.py:17-34 ``amp_aware_dtype``)."""
        t = self.batched_features.batched_tensor
        if torch.is_autocast_enabled() and t.is_floating_point():
            dt = torch.get_autocast_dtype("cuda")
            if t.dtype != dt:
                return t.to(dt)
        return t

    features = feature_tensor
    feats = feature_tensor

    @property
    def nested_features(self) -> Tensor:
        return self.batched_features.to_nested()

    @property
    def padded_features(self):
        return self.batched_features.to_pad()

    def to_pad(self, pad_multiple: Optional[int] = None) -> "Geometry":
        f = self.batched_features
        if f.is_pad and f.pad_multiple == pad_multiple:
            return self
        return self.replace(batched_features=f.to_pad(pad_multiple))

    def to_cat(self) -> "Geometry":
        return self if self.batched_features.is_cat else \
            self.replace(batched_features=self.batched_features.to_cat())

    def sort(self, *args, **kwargs):
        raise NotImplementedError

    def replace_features(self, new_features) -> "Geometry":
        return self.replace(batched_features=new_features)

    @property
    def offsets(self):
        return self.batched_coordinates.offsets

    @property
    def device(self):
        return self.batched_coordinates.device

    @property
    def num_channels(self):
        return self.batched_features.num_channels

    @property
    def batch_size(self) -> int:
        return len(self.offsets) - 1

    @property
    def shape(self):
        return self.batched_features.shape

    @property
    def dtype(self):
        return self.batched_features.dtype

    def _apply(self, fn):
        return self.replace(batched_features=fn(self.batched_features.batched_tensor))

    def _apply_feature_transform(self, feature_transform_fn):
        return self.replace(batched_features=feature_transform_fn(self.feature_tensor))

    def half(self):
        return self._apply(lambda t: t.half())

    def float(self):
        return self._apply(lambda t: t.float())

    def double(self):
        return self._apply(lambda t: t.double())

    def binary_op(self, value, op: str) -> "Geometry":
        """Feature-wise arithmetic with a same-shape geometry, a scalar / one-element tensor or a
        feature-shaped tensor (geometry.py ``binary_op``)."""
        a = self.batched_features.batched_tensor
        if isinstance(value, Geometry):
            assert self.equal_shape(value), f"Shapes do not match. {self} != {value}"
            b = value.batched_features.batched_tensor
        elif isinstance(value, (int, float)) or (torch.is_tensor(value) and value.numel() == 1):
            b = value
        elif isinstance(value, Tensor):
            assert value.shape == a.shape or value.shape == a.shape[-1:], \
                f"tensor operand {tuple(value.shape)} does not match features {tuple(a.shape)}"
            b = value
        else:
            raise NotImplementedError(f"unsupported operand {type(value)}")
        return self.replace(batched_features=getattr(a, op)(b))

    def equal_shape(self, value) -> bool:
        return (self.batched_coordinates.equal_shape(value.batched_coordinates)
                and self.batched_features.equal_shape(value.batched_features))

    def equal_rigorous(self, value) -> bool:
        raise NotImplementedError

    def __add__(self, v):
        return self.binary_op(v, "__add__")

    def __sub__(self, v):
        return self.binary_op(v, "__sub__")

    def __mul__(self, v):
        return self.binary_op(v, "__mul__")

    def __truediv__(self, v):
        return self.binary_op(v, "__truediv__")

    def __floordiv__(self, v):
        return self.binary_op(v, "__floordiv__")

    def __mod__(self, v):
        return self.binary_op(v, "__mod__")

    def __pow__(self, v):
        return self.binary_op(v, "__pow__")

    __radd__ = __add__
    __rmul__ = __mul__

    def __rsub__(self, v):
        return self._apply(lambda t: -t).binary_op(v, "__add__")

    def __rtruediv__(self, v):
        return self._apply(lambda t: t.reciprocal()).binary_op(v, "__mul__")

    def __len__(self) -> int:
        return len(self.batched_coordinates)  # rows, like the reference (geometry.py ``__len__``)

    def numel(self):
        return int(self.offsets[-1]) * self.num_channels

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(offsets={self.offsets.tolist()}, "
                f"feature_shape={tuple(self.batched_features.shape)}, "
                f"coords_shape={tuple(self.batched_coordinates.shape)}, device={self.device}, "
                f"dtype={self.batched_features.dtype})")

    def __str__(self) -> str:
        return (f"{self.__class__.__name__}(feature_shape={tuple(self.batched_features.shape)}, "
                f"coords_shape={tuple(self.batched_coordinates.shape)})")

    @property
    def extra_attributes(self):
        return self._extra_attributes.copy()

    @property
    def cache(self):
        return self._extra_attributes.get("_cache")

    def replace(self, batched_coordinates: Optional[Coords] = None, batched_features=None,
                **kwargs):
        if "_extra_attributes" in kwargs:
            extra = kwargs.pop("_extra_attributes")
            kwargs = {**extra, **kwargs}
        new_coords = batched_coordinates if batched_coordinates is not None \
            else self.batched_coordinates
        new_feats = batched_features if batched_features is not None else self.batched_features
        if isinstance(new_feats, Tensor):
            new_feats = to_batched_features(new_feats, new_coords.offsets)
        return self.__class__(new_coords, new_feats, **{**self._extra_attributes, **kwargs})
