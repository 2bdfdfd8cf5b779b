# SPDX-License-Identifier: Apache-2.0
"""Feature concatenation and pruning of sparse tensors
(drop-in for warpconvnet/nn/functional/sparse_ops.py:13-66 — the skip-connection ``cat`` of the
MinkUNet decoder and the mask-based pruning of generative decoders). Plain torch on the feature /
coordinate tensors; no kernel of this library is involved."""
from __future__ import annotations

import torch
from torch import Tensor

from warpconvnet_b200.geometry.base.geometry import Geometry
from warpconvnet_b200.geometry.types.voxels import Voxels


def cat_spatially_sparse_tensors(*sparse_tensors: Voxels) -> Voxels:
    """Channel-wise concatenation of tensors that share coordinates (same offsets)."""
    offsets = sparse_tensors[0].offsets
    for st in sparse_tensors:
        if len(st.offsets) != len(offsets) or not bool((st.offsets.to(offsets) == offsets).all()):
            raise ValueError("All sparse tensors must have the same offsets")
    feats = torch.cat([st.feature_tensor for st in sparse_tensors], dim=-1)
    return sparse_tensors[0].replace(batched_features=feats)


def prune_spatially_sparse_tensor(spatial_tensor: Geometry, mask: Tensor) -> Geometry:
    """Keep the rows where ``mask`` is true; offsets are recomputed per batch item."""
    n = spatial_tensor.coordinate_tensor.shape[0]
    if mask.shape[0] != n:
        raise ValueError(f"Mask length {mask.shape[0]} must match number of coordinates {n}")
    mask = mask.to(spatial_tensor.device)
    if mask.dtype != torch.bool:
        mask = mask.bool()
    coords = spatial_tensor.batched_coordinates
    if not hasattr(coords, "prune"):
        raise TypeError(f"{coords.__class__.__name__} does not implement prune()")
    attrs = {k: v for k, v in spatial_tensor._extra_attributes.items()
             if k not in ("_cache", "_stride_cache", "_spatial_cache")}  # maps index the old rows
    return spatial_tensor.__class__(coords.prune(mask), spatial_tensor.feature_tensor[mask],
                                    **attrs)
