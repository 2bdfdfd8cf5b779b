# SPDX-License-Identifier: Apache-2.0
"""Batched k-nearest-neighbour search for ``Points``
(behaviour of warpconvnet/geometry/coords/search/knn.py:10-142: per-batch, chunked
``cdist`` + ``topk``; indices are global rows of the reference cloud).
Host-side torch plumbing for now; a grid-hash kNN kernel is the SURVEY.md §8f rank-4 row."""
from __future__ import annotations

import torch
from torch import Tensor

from .search_configs import RealSearchConfig, RealSearchMode
from .search_results import RealSearchResult


@torch.no_grad()
def _knn_one(ref: Tensor, query: Tensor, k: int, chunk: int = 4096) -> Tensor:
    out = []
    for s in range(0, query.shape[0], chunk):
        d = torch.cdist(query[s:s + chunk].float(), ref.float())
        out.append(torch.topk(d, k, dim=1, largest=False, sorted=True).indices)
    return torch.cat(out, dim=0) if out else torch.zeros((0, k), dtype=torch.long, device=ref.device)


@torch.no_grad()
def batched_knn_search(ref: Tensor, ref_offsets: Tensor, query: Tensor, query_offsets: Tensor,
                       k: int) -> Tensor:
    assert len(ref_offsets) == len(query_offsets)
    parts = []
    for b in range(len(ref_offsets) - 1):
        rs, re = int(ref_offsets[b]), int(ref_offsets[b + 1])
        qs, qe = int(query_offsets[b]), int(query_offsets[b + 1])
        assert re - rs >= k, f"batch {b} has fewer than k={k} reference points"
        parts.append(_knn_one(ref[rs:re], query[qs:qe], k) + rs)
    return torch.cat(parts, dim=0)


def neighbor_search(ref, ref_offsets, query, query_offsets, cfg: RealSearchConfig
                    ) -> RealSearchResult:
    if cfg.mode == RealSearchMode.KNN:
        return RealSearchResult(batched_knn_search(ref, ref_offsets, query, query_offsets,
                                                   int(cfg.knn_k)))
    raise NotImplementedError(f"neighbour search mode {cfg.mode} is not part of the hot path")
