# SPDX-License-Identifier: Apache-2.0
"""``Points`` -> ``Voxels`` (drop-in for warpconvnet/geometry/types/conversion/to_voxels.py:4-24 and
the voxel-downsample path behind it, geometry/coords/ops/voxel.py:112-151,
nn/functional/point_pool.py): quantise ``floor(p / voxel_size)``, take the unique voxels per batch
item and reduce the features of the points that fall into each voxel.

Device-side torch glue around the packed 64-bit coordinate key (no host->device syncs; one D2H of
the per-scene voxel counts, which the ``offsets`` contract requires). Voxels come out sorted by
(batch, x, y, z) — deterministic, where the reference's hash-unique order depends on a race — and
``reduction="random"`` keeps the FIRST point of every voxel for the same reason."""
from __future__ import annotations

import torch
from torch import Tensor

from warpconvnet_b200.geometry.coords.integer import IntCoords
from warpconvnet_b200.geometry.coords.ops.batch_index import (batch_index_from_offset,
                                                              offsets_from_batch_index)
from warpconvnet_b200.geometry.coords.ops.stride import pack_sortable, unpack_sortable

_REDUCTIONS = ("random", "mean", "sum", "max", "min")


def _reduce(feats: Tensor, inverse: Tensor, m: int, reduction: str, first_index: Tensor) -> Tensor:
    if reduction == "random":
        return feats[first_index]
    c = feats.shape[1]
    idx = inverse.unsqueeze(1).expand(-1, c)
    if reduction in ("sum", "mean"):
        out = torch.zeros((m, c), dtype=feats.dtype, device=feats.device).index_add_(0, inverse, feats)
        if reduction == "mean":
            counts = torch.bincount(inverse, minlength=m).clamp_min(1).to(feats.dtype)
            out = out / counts.unsqueeze(1)
        return out
    init = torch.zeros((m, c), dtype=feats.dtype, device=feats.device)
    return init.scatter_reduce(0, idx, feats, "amax" if reduction == "max" else "amin",
                               include_self=False)


def points_to_voxels(points, voxel_size: float, reduction="random", unique_method: str = "torch",
                     return_to_unique: bool = False):
    """Voxels whose coordinates are the occupied cells of size ``voxel_size`` and whose features
    reduce the points of each cell. ``return_to_unique=True`` also returns the int64 map
    point -> voxel row."""
    from warpconvnet_b200.geometry.types.voxels import Voxels
    reduction = getattr(reduction, "value", reduction)
    reduction = str(reduction).lower()
    if reduction not in _REDUCTIONS:
        raise ValueError(f"unsupported reduction {reduction!r}; one of {_REDUCTIONS}")
    coords = points.coordinate_tensor
    if not coords.is_cuda:
        raise RuntimeError("warpconvnet_b200 runs on CUDA only; there is no CPU fallback")
    feats = points.feature_tensor
    nb = len(points.offsets) - 1
    with torch.no_grad():
        q = torch.floor(coords / voxel_size).to(torch.int32)
        bidx = batch_index_from_offset(points.offsets, device=coords.device).to(torch.int32)
        keys = pack_sortable(torch.cat([bidx.unsqueeze(1), q], dim=1))
        uniq, inverse = torch.unique(keys, return_inverse=True)       # ascending (batch, x, y, z)
        m = uniq.numel()
        vox_bc = unpack_sortable(uniq)
        offsets = offsets_from_batch_index(vox_bc[:, 0], nb)
        n = coords.shape[0]
        first = torch.full((m,), n, dtype=torch.int64, device=coords.device).scatter_reduce(
            0, inverse, torch.arange(n, device=coords.device), "amin", include_self=True)
    out_feats = _reduce(feats, inverse, m, reduction, first)
    vox = Voxels(IntCoords(vox_bc[:, 1:].contiguous(), offsets=offsets, voxel_size=voxel_size),
                 out_feats, voxel_size=voxel_size)
    return (vox, inverse) if return_to_unique else vox
