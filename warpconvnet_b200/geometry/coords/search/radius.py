# SPDX-License-Identifier: Apache-2.0
"""Radius neighbour search for ``Points`` (drop-in for
warpconvnet/geometry/coords/search/radius.py:162-291: same signatures, same return triple
``(neighbor_index, neighbor_distance, neighbor_split)``). The reference loops over batch items
in Python and runs either its Warp hash-grid kernels or chunked ``torch.cdist``; here ONE pair of
device passes over the uniform grid of ``csrc/knn.cu`` (count, then fill) handles all batch items.
Order of the neighbours inside a row is unspecified, as in the reference. No CPU fallback."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from warpconvnet_b200 import _ops


def _check(t: Tensor):
    if not t.is_cuda:
        raise RuntimeError("warpconvnet_b200 neighbour search runs on CUDA only (no CPU fallback)")


@torch.no_grad()
def radius_search(points: Tensor, queries: Tensor, radius: float,
                  grid_dim=None) -> Tuple[Tensor, Tensor, Tensor]:
    """Single cloud: (int32 [Q] indices into ``points``, float32 [Q] distances, int32 [M + 1]
    splits); ``grid_dim`` is accepted and ignored like in the reference."""
    _check(points)
    n, m = points.shape[0], queries.shape[0]
    if n == 0 or m == 0:
        dev = queries.device
        return (torch.zeros(0, dtype=torch.int32, device=dev),
                torch.zeros(0, dtype=torch.float32, device=dev),
                torch.zeros(m + 1, dtype=torch.int32, device=dev))
    ro = torch.tensor([0, n], dtype=torch.int32)
    qo = torch.tensor([0, m], dtype=torch.int32)
    idx, dist, splits = _ops.radius_search(points.float(), ro, queries.float(), qo, float(radius))
    return idx, dist, splits.int()


@torch.no_grad()
def batched_radius_search(ref_positions: Tensor, ref_offsets: Tensor, query_positions: Tensor,
                          query_offsets: Tensor, radius: float,
                          grid_dim: Optional[int] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """(int64 [Q] GLOBAL reference rows, float32 [Q] distances, int64 [M + 1] splits)."""
    _check(ref_positions)
    assert len(ref_offsets) == len(query_offsets)
    assert int(ref_offsets[-1]) == ref_positions.shape[0], \
        f"Last offset {int(ref_offsets[-1])} != {ref_positions.shape[0]}"
    assert int(query_offsets[-1]) == query_positions.shape[0], \
        f"Last offset {int(query_offsets[-1])} != {query_positions.shape[0]}"
    idx, dist, splits = _ops.radius_search(ref_positions.float(), ref_offsets,
                                           query_positions.float(), query_offsets, float(radius))
    return idx.long(), dist, splits
