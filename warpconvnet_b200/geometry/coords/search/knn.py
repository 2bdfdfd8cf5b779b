# SPDX-License-Identifier: Apache-2.0
"""Batched k-nearest-neighbour search for ``Points``
(drop-in for warpconvnet/geometry/coords/search/knn.py:108-142: indices are global rows of the
reference cloud, searched per batch item). The reference loops over batch items in Python and
runs chunked ``torch.cdist`` + ``torch.topk`` (O(M*N)); here ONE call of the device grid-kNN
kernel (``csrc/knn.cu``) handles all batch items. There is no CPU / library fallback."""
from __future__ import annotations

import torch
from torch import Tensor

from warpconvnet_b200 import _ops

from .search_configs import RealSearchConfig, RealSearchMode
from .search_results import RealSearchResult


@torch.no_grad()
def batched_knn_search(ref_positions: Tensor, ref_offsets: Tensor, query_positions: Tensor,
                       query_offsets: Tensor, k: int, search_method: str = "grid",
                       chunk_size: int = 4096) -> Tensor:
    """int64 [M, k]: k nearest reference rows of every query, ascending distance."""
    assert len(ref_offsets) == len(query_offsets)
    counts = (ref_offsets[1:] - ref_offsets[:-1])
    assert 0 < k <= 64, f"k must be in [1, 64], got {k}"
    assert int(counts.min()) >= k, \
        f"k must not exceed the number of reference points of any batch item. K: {k}"
    if not ref_positions.is_cuda:
        raise RuntimeError("warpconvnet_b200 neighbour search runs on CUDA only (no CPU fallback)")
    return _ops.knn_search(ref_positions.float(), ref_offsets, query_positions.float(),
                           query_offsets, int(k))


def neighbor_search(ref, ref_offsets, query, query_offsets, cfg: RealSearchConfig
                    ) -> RealSearchResult:
    if cfg.mode == RealSearchMode.KNN:
        return RealSearchResult(batched_knn_search(ref, ref_offsets, query, query_offsets,
                                                   int(cfg.knn_k)))
    if cfg.mode == RealSearchMode.RADIUS:
        from .radius import batched_radius_search
        assert cfg.radius is not None, "RealSearchConfig(mode='radius') needs a radius"
        idx, dist, splits = batched_radius_search(ref, ref_offsets, query, query_offsets,
                                                  float(cfg.radius), cfg.grid_dim)
        res = RealSearchResult(idx, splits)
        res.neighbor_distances = dist
        return res
    raise NotImplementedError(f"neighbour search mode {cfg.mode} is not part of the hot path "
                              "(SURVEY.md §2b: voxel-grid search is out of scope)")
