# SPDX-License-Identifier: Apache-2.0
"""Strided (downsampled) output coordinates
(drop-in for warpconvnet/geometry/coords/ops/stride.py:18-56).

The reference floor-divides through float, dedups with a racy hash insert and then argsorts the
batch column (row order is nondeterministic). Here the unique set is taken on the packed 64-bit
coordinate key, so rows come out sorted by (batch, x, y, z) — deterministic, already
batch-contiguous, and spatially coherent for the next level's gathers.
"""
from typing import Tuple

import torch
from torch import Tensor

from warpconvnet_b200.geometry.coords.ops.batch_index import offsets_from_batch_index
from warpconvnet_b200.utils.ntuple import ntuple

_OFF = 1 << 17  # bias that maps [-131072, 131071] to [0, 262143]


def pack_sortable(bcoords: Tensor) -> Tensor:
    """int64 key, monotone in (batch, x, y, z) lexicographic order."""
    c = bcoords.long()
    return (c[:, 0] << 54) | ((c[:, 1] + _OFF) << 36) | ((c[:, 2] + _OFF) << 18) | (c[:, 3] + _OFF)


def unpack_sortable(keys: Tensor) -> Tensor:
    b = keys >> 54
    x = ((keys >> 36) & 0x3FFFF) - _OFF
    y = ((keys >> 18) & 0x3FFFF) - _OFF
    z = (keys & 0x3FFFF) - _OFF
    return torch.stack([b, x, y, z], dim=1).int()


@torch.no_grad()
def unique_coords(bcoords: Tensor) -> Tuple[Tensor, Tensor]:
    """(unique rows sorted by key, index of the first occurrence of each)."""
    pad = bcoords.shape[1] == 3
    c = torch.nn.functional.pad(bcoords, (0, 1), value=0) if pad else bcoords
    keys = pack_sortable(c)
    order = torch.argsort(keys, stable=True)
    sk = keys[order]
    first = torch.ones_like(sk, dtype=torch.bool)
    first[1:] = sk[1:] != sk[:-1]
    idx = order[first]
    out = unpack_sortable(sk[first])
    return (out[:, :3] if pad else out), idx


@torch.no_grad()
def stride_coords(batch_indexed_coords: Tensor, stride: Tuple[int, ...], order=None
                  ) -> Tuple[Tensor, Tensor]:
    num_spatial_dims = batch_indexed_coords.shape[1] - 1
    stride = ntuple(stride, ndim=num_spatial_dims)
    if all(s == 1 for s in stride):
        return batch_indexed_coords, offsets_from_batch_index(batch_indexed_coords[:, 0])
    # floor division by Python scalars: a torch.tensor(stride, device=cuda) here would be a
    # synchronous pageable H2D copy, i.e. a host sync in every strided layer
    discretized = batch_indexed_coords.clone()
    if len(set(stride)) == 1:
        discretized[:, 1:] = torch.div(batch_indexed_coords[:, 1:], stride[0],
                                       rounding_mode="floor")
    else:
        for d, s in enumerate(stride):
            discretized[:, d + 1] = torch.div(batch_indexed_coords[:, d + 1], s,
                                              rounding_mode="floor")
    unique, _ = unique_coords(discretized)
    return unique.contiguous(), offsets_from_batch_index(unique[:, 0])
