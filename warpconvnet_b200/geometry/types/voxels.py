# SPDX-License-Identifier: Apache-2.0
"""Sparse voxel container (drop-in for warpconvnet/geometry/types/voxels.py:23-317, the subset the
sparse-conv path uses: constructors, ``replace``, ``unique``, ``tensor_stride``, ``cache``,
``batch_indexed_coordinates``, ``to_dense``)."""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor

from warpconvnet_b200.geometry.base.batched import CatFeatures, Features, to_batched_features
from warpconvnet_b200.geometry.base.geometry import Geometry
from warpconvnet_b200.geometry.coords.integer import IntCoords
from warpconvnet_b200.geometry.coords.ops.batch_index import offsets_from_batch_index


class Voxels(Geometry):
    def __init__(self, batched_coordinates, batched_features, offsets: Optional[Tensor] = None,
                 device: Optional[str] = None, **kwargs):
        tensor_stride = kwargs.pop("tensor_stride", None) or kwargs.pop("stride", None)
        kwargs.pop("stride", None)
        if isinstance(batched_coordinates, (list, tuple)):
            assert isinstance(batched_features, (list, tuple)), \
                "If coords is a list, features must be a list too."
            assert len(batched_coordinates) == len(batched_features)
            assert all(len(c) == len(f) for c, f in zip(batched_coordinates, batched_features))
            batched_coordinates = IntCoords(list(batched_coordinates), device=device,
                                            tensor_stride=tensor_stride)
        elif isinstance(batched_coordinates, Tensor):
            assert isinstance(batched_features, Tensor) and offsets is not None, \
                "If coordinate is a tensor, features must be a tensor and offsets must be provided."
            batched_coordinates = IntCoords(batched_coordinates, offsets=offsets, device=device,
                                            tensor_stride=tensor_stride)
        else:
            if tensor_stride is not None:
                batched_coordinates.set_tensor_stride(tensor_stride)
        if isinstance(batched_features, (list, tuple)):
            batched_features = CatFeatures(list(batched_features), device=device)
        elif isinstance(batched_features, Tensor):
            batched_features = to_batched_features(batched_features, batched_coordinates.offsets,
                                                   device=device)
        Geometry.__init__(self, batched_coordinates, batched_features, **kwargs)

    def unique(self) -> "Voxels":
        """One voxel per distinct coordinate (voxels.py:271-278); keeps the first occurrence,
        rows sorted by (batch, x, y, z)."""
        from warpconvnet_b200.geometry.coords.ops.stride import unique_coords
        uniq, idx = unique_coords(self.batch_indexed_coordinates)
        offs = offsets_from_batch_index(uniq[:, 0], self.batch_size)
        coords = IntCoords(uniq[:, 1:].contiguous(), offs, tensor_stride=self.tensor_stride)
        feats = CatFeatures(self.batched_features.batched_tensor[idx], offs)
        attrs = {k: v for k, v in self._extra_attributes.items() if k != "_cache"}
        return self.__class__(coords, feats, **attrs)

    @property
    def coordinate_hashmap(self):
        return self.batched_coordinates.hashmap

    @property
    def voxel_size(self):
        return self._extra_attributes.get("voxel_size", None)

    @property
    def ordering(self):
        return self._extra_attributes.get("ordering", None)

    @property
    def tensor_stride(self):
        return self.batched_coordinates.tensor_stride

    stride = tensor_stride

    def set_tensor_stride(self, tensor_stride):
        self.batched_coordinates.set_tensor_stride(tensor_stride)

    @property
    def batch_indexed_coordinates(self) -> Tensor:
        return self.batched_coordinates.batch_indexed_coordinates

    @property
    def spatial_cache(self) -> dict:
        return self._extra_attributes.setdefault("_spatial_cache", {})

    def to_dense(self, channel_dim: int = 1, spatial_shape=None, min_coords=None) -> Tensor:
        """Dense [B, C, X, Y, Z] tensor (voxels.py:139-228, default channel-first layout)."""
        bc = self.batch_indexed_coordinates.long()
        sp = bc[:, 1:]
        if min_coords is None:
            min_coords = sp.min(dim=0).values
        else:
            min_coords = torch.as_tensor(min_coords, device=sp.device)
        sp = sp - min_coords
        if spatial_shape is None:
            spatial_shape = tuple(int(v) + 1 for v in sp.max(dim=0).values.tolist())
        feats = self.batched_features.batched_tensor
        dense = torch.zeros((self.batch_size, *spatial_shape, feats.shape[1]), dtype=feats.dtype,
                            device=feats.device)
        dense[(bc[:, 0], *[sp[:, i] for i in range(sp.shape[1])])] = feats
        if channel_dim == 1:
            perm = (0, dense.dim() - 1, *range(1, dense.dim() - 1))
            dense = dense.permute(*perm).contiguous()
        return dense
