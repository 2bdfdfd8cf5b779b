# SPDX-License-Identifier: Apache-2.0
"""Neighbour-search configuration for ``Points`` (search_configs.py in the reference)."""
from dataclasses import dataclass
from enum import Enum
from typing import Optional


class RealSearchMode(Enum):
    RADIUS = "radius"
    KNN = "knn"
    VOXEL = "voxel"


@dataclass(frozen=True)
class RealSearchConfig:
    mode: RealSearchMode = RealSearchMode.KNN
    radius: Optional[float] = None
    knn_k: Optional[int] = None
    grid_dim: Optional[int] = None

    def __init__(self, mode="knn", radius=None, knn_k=None, grid_dim=None, **kwargs):
        if isinstance(mode, str):
            mode = RealSearchMode(mode.lower())
        object.__setattr__(self, "mode", mode)
        object.__setattr__(self, "radius", radius)
        object.__setattr__(self, "knn_k", knn_k)
        object.__setattr__(self, "grid_dim", grid_dim)

    def replace(self, **kw):
        d = dict(mode=self.mode, radius=self.radius, knn_k=self.knn_k, grid_dim=self.grid_dim)
        d.update(kw)
        return RealSearchConfig(**d)
