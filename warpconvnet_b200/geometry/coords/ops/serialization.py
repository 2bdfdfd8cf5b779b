# SPDX-License-Identifier: Apache-2.0
"""Z-order (Morton) serialisation of integer coordinates — the ordering behind ``Voxels.sort`` /
``Points.sort`` (warpconvnet/geometry/coords/ops/serialization.py:22-260, csrc/morton_code.cu:14-54).

Same code definition as the reference: coordinates are shifted to start at 0, permuted by the
ordering (``MORTON_XZY`` reads the axes as x, z, y ...), and the bits of the three axes are
interleaved with the FIRST permuted axis in the lowest bit. 21 bits per axis; the batch index is
not packed into the code (the reference's batched path keeps 16 bits per axis for that) — rows are
ordered by (batch, code) with two stable sorts. Device-side torch integer ops; outside the hot path.
"""
from __future__ import annotations

from enum import Enum
from typing import NamedTuple, Optional

import torch
from torch import Tensor


class POINT_ORDERING(Enum):
    RANDOM = 0
    MORTON_XYZ = 1
    MORTON_XZY = 2
    MORTON_YXZ = 3
    MORTON_YZX = 4
    MORTON_ZXY = 5
    MORTON_ZYX = 6


STR2POINT_ORDERING = {"random": POINT_ORDERING.RANDOM, "morton": POINT_ORDERING.MORTON_XYZ,
                      **{f"morton_{''.join(p)}": POINT_ORDERING[f"MORTON_{''.join(p).upper()}"]
                         for p in ("xyz", "xzy", "yxz", "yzx", "zxy", "zyx")}}
_AXES = {o: ["xyz".index(ch) for ch in o.name[-3:].lower()]
         for o in POINT_ORDERING if o is not POINT_ORDERING.RANDOM}


class SerializationResult(NamedTuple):
    codes: Tensor
    perm: Optional[Tensor] = None
    inverse_perm: Optional[Tensor] = None


def _spread3(v: Tensor) -> Tensor:
    """21-bit value -> every bit moved to position 3*i."""
    v = v & 0x1FFFFF
    v = (v | (v << 32)) & 0x1F00000000FFFF
    v = (v | (v << 16)) & 0x1F0000FF0000FF
    v = (v | (v << 8)) & 0x100F00F00F00F00F
    v = (v | (v << 4)) & 0x10C30C30C30C30C3
    v = (v | (v << 2)) & 0x1249249249249249
    return v


@torch.no_grad()
def morton_code(coords: Tensor, order=POINT_ORDERING.MORTON_XYZ) -> Tensor:
    """int64 z-order code of ``[N, 3]`` (or batch-indexed ``[N, 4]``: the batch column is ignored)
    integer coordinates."""
    if isinstance(order, str):
        order = STR2POINT_ORDERING[order]
    assert order in _AXES, f"Order '{order}' not supported for morton code"
    c = coords[:, -3:].long()
    if c.shape[0] == 0:
        return torch.empty(0, dtype=torch.int64, device=coords.device)
    c = (c - c.min(dim=0).values)[:, _AXES[order]]
    return _spread3(c[:, 0]) | (_spread3(c[:, 1]) << 1) | (_spread3(c[:, 2]) << 2)


@torch.no_grad()
def encode(grid_coord: Tensor, batch_offsets: Optional[Tensor] = None,
           order=POINT_ORDERING.MORTON_XYZ, return_perm: bool = False,
           return_inverse: bool = False) -> SerializationResult:
    """Codes (and the permutation that sorts every batch item by its code)."""
    if isinstance(order, str):
        order = STR2POINT_ORDERING[order]
    n = grid_coord.shape[0]
    if order is POINT_ORDERING.RANDOM:
        codes = torch.rand(n, device=grid_coord.device)
    else:
        codes = morton_code(grid_coord, order)
    if not (return_perm or return_inverse):
        return SerializationResult(codes)
    perm = torch.argsort(codes, stable=True)
    if batch_offsets is not None and len(batch_offsets) > 2:
        from .batch_index import batch_index_from_offset
        bidx = batch_index_from_offset(batch_offsets, device=grid_coord.device)
        perm = perm[torch.argsort(bidx[perm], stable=True)]
    inverse = None
    if return_inverse:
        inverse = torch.empty_like(perm)
        inverse[perm] = torch.arange(n, device=perm.device)
    return SerializationResult(codes, perm, inverse)
