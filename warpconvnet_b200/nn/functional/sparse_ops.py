# SPDX-License-Identifier: Apache-2.0
"""Channel concatenation and row pruning for sparse tensors — the skip-connection join of the
MinkUNet decoder and the keep-mask step of generative decoders. Same call signatures and
exception types as warpconvnet/nn/functional/sparse_ops.py:13-66; torch indexing only, none of
this library's kernels is involved."""
from __future__ import annotations

import torch
from torch import Tensor

from warpconvnet_b200.geometry.base.geometry import _ROW_CACHES, Geometry
from warpconvnet_b200.geometry.types.voxels import Voxels

# kernel maps / strided coordinate sets cached on a tensor index its OLD rows
_ROW_INDEXED_ATTRS = _ROW_CACHES


def cat_spatially_sparse_tensors(*sparse_tensors: Voxels) -> Voxels:
    """``[N, C1 + C2 + ...]`` features on the coordinates of the first argument; every argument
    must describe the same rows (equal per-scene offsets)."""
    head, *others = sparse_tensors
    layout = head.offsets.tolist()
    if any(t.offsets.tolist() != layout for t in others):
        raise ValueError(f"cannot concatenate sparse tensors with different offsets "
                         f"({[t.offsets.tolist() for t in sparse_tensors]})")
    columns = [t.feature_tensor for t in sparse_tensors]
    return head.replace(batched_features=torch.cat(columns, dim=-1))


def prune_spatially_sparse_tensor(spatial_tensor: Geometry, mask: Tensor) -> Geometry:
    """The rows of ``spatial_tensor`` whose ``mask`` entry is true (any dtype, any device);
    per-scene offsets are recounted and row-indexed caches are dropped."""
    rows = spatial_tensor.coordinate_tensor.shape[0]
    if mask.shape[0] != rows:
        raise ValueError(f"mask has {mask.shape[0]} entries for {rows} coordinates")
    coords = spatial_tensor.batched_coordinates
    if not callable(getattr(coords, "prune", None)):
        raise TypeError(f"{type(coords).__name__} coordinates cannot be pruned")
    keep = mask.to(device=spatial_tensor.device, dtype=torch.bool)
    carried = {k: v for k, v in spatial_tensor._extra_attributes.items()
               if k not in _ROW_INDEXED_ATTRS}
    return type(spatial_tensor)(coords.prune(keep), spatial_tensor.feature_tensor[keep], **carried)
