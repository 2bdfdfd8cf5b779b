# SPDX-License-Identifier: Apache-2.0
"""Point-cloud container (drop-in for warpconvnet/geometry/types/points.py:33-326, the subset
``PointConv`` needs: constructors, ``replace``, ``neighbors`` with a per-instance cache)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from warpconvnet_b200.geometry.base.batched import CatFeatures, to_batched_features
from warpconvnet_b200.geometry.base.geometry import Geometry
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
