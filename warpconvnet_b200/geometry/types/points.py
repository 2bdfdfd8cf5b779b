# SPDX-License-Identifier: Apache-2.0
"""Point-cloud container (drop-in for warpconvnet/geometry/types/points.py:33-326, the subset
``PointConv`` needs: constructors, ``replace``, ``neighbors`` with a per-instance cache)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from warpconvnet_b200.geometry.base.batched import CatFeatures, to_batched_features
from warpconvnet_b200.geometry.base.geometry import _ROW_CACHES, Geometry
from warpconvnet_b200.geometry.coords.integer import RealCoords
from warpconvnet_b200.geometry.coords.search.cache import RealSearchCache
from warpconvnet_b200.geometry.coords.search.search_configs import RealSearchConfig


class Points(Geometry):
    def __init__(self, batched_coordinates, batched_features, offsets: Optional[Tensor] = None,
                 device: Optional[str] = None, **kwargs):
        if isinstance(batched_coordinates, (list, tuple)):
            assert isinstance(batched_features, (list, tuple)), \
                "If coords is a list, features must be a list too."
            assert len(batched_coordinates) == len(batched_features)
            assert all(len(c) == len(f) for c, f in zip(batched_coordinates, batched_features))
            batched_coordinates = RealCoords(list(batched_coordinates), device=device)
        elif isinstance(batched_coordinates, Tensor):
            assert isinstance(batched_features, Tensor) and offsets is not None, \
                "If coordinate is a tensor, features must be a tensor and offsets must be provided."
            batched_coordinates = RealCoords(batched_coordinates, offsets=offsets, device=device)
        if isinstance(batched_features, (list, tuple)):
            batched_features = CatFeatures(list(batched_features), device=device)
        elif isinstance(batched_features, Tensor):
            batched_features = to_batched_features(batched_features, batched_coordinates.offsets,
                                                   device=device)
        Geometry.__init__(self, batched_coordinates, batched_features, **kwargs)

    @classmethod
    def from_list_of_coordinates(cls, coordinates, features=None, encoding_channels=None,
                                 encoding_range=None, encoding_dim: int = -1) -> "Points":
        """Points from a list of ``[N_i, 3]`` tensors; without features a sinusoidal encoding of
        the coordinates is used (points.py:283-316)."""
        if isinstance(coordinates, Tensor):
            coordinates = list(coordinates)
        if features is None:
            assert encoding_range is not None, \
                "Encoding range must be provided if encoding channels are provided"
            from warpconvnet_b200.nn.encodings import sinusoidal_encoding
            features = [sinusoidal_encoding(c, encoding_channels, encoding_range, encoding_dim)
                        for c in coordinates]
        return cls(RealCoords(list(coordinates)), CatFeatures(list(features)))

    def _take(self, rows: Tensor, offsets: Tensor, **extra) -> "Points":
        attrs = {k: v for k, v in self._extra_attributes.items() if k not in _ROW_CACHES}
        attrs.update(extra)
        return self.__class__(RealCoords(self.coordinate_tensor[rows], offsets),
                              CatFeatures(self.feature_tensor[rows], offsets), **attrs)

    def sort(self, voxel_size: float, ordering=None) -> "Points":
        """Z-order of the points' ``floor(p / voxel_size)`` cells inside every batch item
        (points.py:92-120); points of one cell keep their input order."""
        from warpconvnet_b200.geometry.coords.ops.serialization import POINT_ORDERING, encode
        ordering = POINT_ORDERING.MORTON_XYZ if ordering is None else ordering
        res = encode(torch.floor(self.coordinate_tensor / voxel_size).int(),
                     batch_offsets=self.offsets, order=ordering, return_perm=True)
        return self._take(res.perm, self.offsets)

    def voxel_downsample(self, voxel_size: float, reduction="random") -> "Points":
        """One point per occupied ``voxel_size`` cell (points.py:122-187): the first point of the
        cell, with its own features (``random``) or the cell's reduced features. Cells come out
        sorted by (batch, x, y, z)."""
        from warpconvnet_b200.geometry.coords.ops.batch_index import (batch_index_from_offset,
                                                                      offsets_from_batch_index)
        from warpconvnet_b200.geometry.coords.ops.stride import pack_sortable
        from warpconvnet_b200.geometry.types.conversion.to_voxels import _reduce
        reduction = str(getattr(reduction, "value", reduction)).lower()
        coords = self.coordinate_tensor
        n = coords.shape[0]
        with torch.no_grad():
            cell = torch.floor(coords / voxel_size).to(torch.int32)
            bidx = batch_index_from_offset(self.offsets, device=coords.device).to(torch.int32)
            uniq, inverse = torch.unique(pack_sortable(torch.cat([bidx[:, None], cell], dim=1)),
                                         return_inverse=True)
            m = uniq.numel()
            first = torch.full((m,), n, dtype=torch.int64, device=coords.device).scatter_reduce(
                0, inverse, torch.arange(n, device=coords.device), "amin", include_self=True)
            offsets = offsets_from_batch_index(uniq >> 54, self.batch_size)
        out = self._take(first, offsets, voxel_size=voxel_size)
        if reduction != "random":
            out = out.replace(batched_features=_reduce(self.feature_tensor, inverse, m, reduction,
                                                       first))
        return out

    def random_downsample(self, num_sample_points: int) -> "Points":
        """At most ``num_sample_points`` points of every batch item, without replacement
        (points.py:189-208)."""
        rows, counts = [], [0]
        for b in range(self.batch_size):
            lo, hi = int(self.offsets[b]), int(self.offsets[b + 1])
            k = min(num_sample_points, hi - lo)
            rows.append(lo + torch.randperm(hi - lo, device=self.device)[:k])
            counts.append(counts[-1] + k)
        return self._take(torch.cat(rows), torch.LongTensor(counts))

    def contiguous(self) -> "Points":
        if self.coordinate_tensor.is_contiguous() and \
                self.batched_features.batched_tensor.is_contiguous():
            return self
        return self.replace(
            batched_coordinates=RealCoords(self.coordinate_tensor.contiguous(), self.offsets),
            batched_features=self.batched_features.batched_tensor.contiguous())

    @property
    def ordering(self):
        return self._extra_attributes.get("ordering", None)

    def neighbors(self, search_args: RealSearchConfig, query_coords: Optional[RealCoords] = None):
        """Neighbour search against this cloud, cached by (config, offsets)
        (points.py:237-272)."""
        from warpconvnet_b200.geometry.coords.search.knn import neighbor_search
        if query_coords is None:
            query_coords = self.batched_coordinates
        cache = self._extra_attributes.get("_cache")
        if not isinstance(cache, RealSearchCache):
            cache = RealSearchCache()
            self._extra_attributes["_cache"] = cache
        hit = cache.get(search_args, self.offsets, query_coords.offsets)
        if hit is not None:
            return hit
        result = neighbor_search(self.coordinate_tensor, self.offsets,
                                 query_coords.batched_tensor, query_coords.offsets, search_args)
        cache.put(search_args, self.offsets, query_coords.offsets, result)
        return result

    def to_voxels(self, voxel_size: float, reduction="mean"):
        """Quantise to voxels of size ``voxel_size`` (points.py:318-326)."""
        from warpconvnet_b200.geometry.types.conversion.to_voxels import points_to_voxels
        return points_to_voxels(self, voxel_size, reduction)

    @property
    def voxel_size(self):
        return self._extra_attributes.get("voxel_size", None)
