# SPDX-License-Identifier: Apache-2.0
"""Sparse voxel container (drop-in for warpconvnet/geometry/types/voxels.py:23-317, the subset the
sparse-conv path uses: constructors, ``replace``, ``unique``, ``tensor_stride``, ``cache``,
``batch_indexed_coordinates``, ``to_dense``)."""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor

from warpconvnet_b200.geometry.base.batched import CatFeatures, Features, to_batched_features
from warpconvnet_b200.geometry.base.geometry import _ROW_CACHES, Geometry
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

    @classmethod
    def from_dense(cls, dense_tensor: Tensor, dense_tensor_channel_dim: int = 1,
                   target_spatial_sparse_tensor: Optional["Voxels"] = None,
                   dense_max_coords=None, **kwargs) -> "Voxels":
        """Voxels of the non-zero cells of a dense ``[B, C, *spatial]`` tensor, or — with a target —
        the dense values at the target's coordinates (voxels.py:50-106)."""
        if dense_tensor_channel_dim != -1 and dense_tensor_channel_dim != dense_tensor.ndim - 1:
            dense_tensor = dense_tensor.moveaxis(dense_tensor_channel_dim, -1)
        flat = dense_tensor.flatten(0, -2)
        dims = dense_tensor.shape[:-1]

        def ravel(bc: Tensor) -> Tensor:
            idx = bc[:, 0].long()
            for d in range(1, len(dims)):
                idx = idx * dims[d] + bc[:, d].long()
            return idx

        if target_spatial_sparse_tensor is None:
            cells = torch.nonzero(dense_tensor.abs().sum(dim=-1)).int()   # sorted by (b, x, y, z)
            offs = offsets_from_batch_index(cells[:, 0], dense_tensor.shape[0])
            return cls(cells[:, 1:].contiguous(), flat[ravel(cells)], offsets=offs, **kwargs)
        target = target_spatial_sparse_tensor
        assert target.num_spatial_dims == dense_tensor.ndim - 2
        assert target.batch_size == dense_tensor.shape[0]
        bc = target.batch_indexed_coordinates
        if dense_max_coords is not None:
            assert bool((bc[:, 1:].max(dim=0).values.cpu() <= torch.as_tensor(dense_max_coords)).all())
        return target.replace(batched_features=flat[ravel(bc)])

    def to_point(self, voxel_size: Optional[float] = None):
        """Points at ``coordinate * voxel_size (* tensor_stride)`` (voxels.py:230-246)."""
        from warpconvnet_b200.geometry.coords.integer import RealCoords
        from warpconvnet_b200.geometry.types.points import Points
        if voxel_size is None:
            assert self.voxel_size is not None, \
                "Voxel size must be provided or the object must have been initialized with one"
            voxel_size = self.voxel_size
        scale = self.coordinate_tensor.new_tensor(
            [voxel_size * s for s in (self.tensor_stride or (1,) * self.num_spatial_dims)],
            dtype=torch.float32)
        return Points(RealCoords(self.coordinate_tensor.float() * scale, self.offsets),
                      self.batched_features)

    def sort(self, ordering=None) -> "Voxels":
        """Rows of every batch item in z-order (voxels.py:248-270)."""
        from warpconvnet_b200.geometry.coords.ops.serialization import POINT_ORDERING, encode
        ordering = POINT_ORDERING.MORTON_XYZ if ordering is None else ordering
        if ordering == self.ordering:
            return self
        res = encode(self.coordinate_tensor, batch_offsets=self.offsets, order=ordering,
                     return_perm=True)
        attrs = {k: v for k, v in self._extra_attributes.items() if k not in _ROW_CACHES}
        attrs.update(ordering=ordering, code=res.codes)
        coords = IntCoords(self.coordinate_tensor[res.perm], self.offsets,
                           voxel_size=self.batched_coordinates.voxel_size,
                           tensor_stride=self.tensor_stride)
        feats = CatFeatures(self.batched_features.batched_tensor[res.perm], self.offsets)
        return self.__class__(coords, feats, **attrs)

    def unique(self) -> "Voxels":
        """One voxel per distinct coordinate (voxels.py:271-278); keeps the first occurrence,
        rows sorted by (batch, x, y, z)."""
        from warpconvnet_b200.geometry.coords.ops.stride import unique_with_offsets
        uniq, idx, offs = unique_with_offsets(self.batch_indexed_coordinates, self.batch_size)
        coords = IntCoords(uniq[:, 1:].contiguous(), offs, tensor_stride=self.tensor_stride)
        feats = CatFeatures(self.batched_features.batched_tensor[idx], offs)
        attrs = {k: v for k, v in self._extra_attributes.items() if k not in _ROW_CACHES}
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
