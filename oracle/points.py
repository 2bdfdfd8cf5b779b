# SPDX-License-Identifier: Apache-2.0
"""CPU oracle for the ``Points`` / ``PointConv`` row of the path.  TEST INFRASTRUCTURE ONLY.

* ``knn`` — brute-force k nearest neighbours per batch item in float64: what the reference's
  ``batched_knn_search`` computes (geometry/coords/search/knn.py:108-142, ``cdist`` + ``topk``),
  without its chunking. Ties are broken by the smaller index (``torch.topk`` leaves the order of
  equal distances unspecified).
* ``point_conv_forward`` — restatement of ``PointConv.forward`` (nn/modules/point_conv.py:231-282)
  on CPU tensors for a module whose sub-MLPs are given: gather edge features, edge MLP, row
  reduction over each query's neighbours, output MLP.

PARITY UNPINNED for this row: the reference's PointConv tests assert shapes and gradient existence
only (tests/nn/test_point_conv.py:26-105), its reduction lives in the absent third-party
``torch_scatter`` (un-pinned version) and kNN tie order is unspecified, so there is no golden
vector to pin against; the oracle follows the reference source line by line instead.
"""
from __future__ import annotations

import numpy as np
import torch


def knn(ref: np.ndarray, ref_offsets, query: np.ndarray, query_offsets, k: int):
    """(idx int64 [M, k] global rows, dist float64 [M, k]) ascending distance."""
    ref = np.asarray(ref, np.float64)
    query = np.asarray(query, np.float64)
    idx_parts, d_parts = [], []
    for b in range(len(ref_offsets) - 1):
        rs, re = int(ref_offsets[b]), int(ref_offsets[b + 1])
        qs, qe = int(query_offsets[b]), int(query_offsets[b + 1])
        r, q = ref[rs:re], query[qs:qe]
        d2 = ((q[:, None, :] - r[None, :, :]) ** 2).sum(-1)
        order = np.lexsort((np.broadcast_to(np.arange(re - rs), d2.shape), d2), axis=1)[:, :k]
        idx_parts.append(order + rs)
        d_parts.append(np.sqrt(np.take_along_axis(d2, order, axis=1)))
    return np.concatenate(idx_parts).astype(np.int64), np.concatenate(d_parts)


def point_conv_forward(in_feats, query_feats, in_coords, query_coords, neighbor_idx, k,
                       edge_mlp, out_mlp, reductions=("mean",), use_rel_pos=False,
                       pos_encoding=None):
    """All tensors CPU, same dtype as the (CPU) sub-modules."""
    idx = torch.as_tensor(neighbor_idx).long().reshape(-1)
    m = query_feats.shape[0]
    edge = [in_feats[idx], query_feats.repeat_interleave(k, dim=0)]
    if use_rel_pos or pos_encoding is not None:
        rel = in_coords[idx] - query_coords.repeat_interleave(k, dim=0)
        edge.append(pos_encoding(rel) if pos_encoding is not None else rel)
    e = edge_mlp(torch.cat(edge, dim=1)).view(m, k, -1)
    outs = []
    for r in reductions:
        if r == "mean":
            outs.append(e.mean(1))
        elif r == "sum":
            outs.append(e.sum(1))
        elif r == "max":
            outs.append(e.max(1).values)
        elif r == "min":
            outs.append(e.min(1).values)
        else:
            raise ValueError(r)
    return out_mlp(torch.cat(outs, dim=-1))


def radius(ref: np.ndarray, ref_offsets, query: np.ndarray, query_offsets, r: float):
    """Brute-force radius search per batch item in float64 — what the reference's
    ``batched_radius_search`` computes (geometry/coords/search/radius.py:127-158,228-291:
    ``cdist`` then ``dists <= radius``). Returns (indices int64 [Q] of global reference rows,
    ascending inside every row; distances float64 [Q]; row_splits int64 [M + 1]).
    Pinned by tests/golden/radius_*.npz (outputs of the reference's own ``radius_search``)."""
    ref = np.asarray(ref, np.float64)
    query = np.asarray(query, np.float64)
    idx_parts, d_parts, counts = [], [], []
    for b in range(len(ref_offsets) - 1):
        rs, re = int(ref_offsets[b]), int(ref_offsets[b + 1])
        qs, qe = int(query_offsets[b]), int(query_offsets[b + 1])
        rr, q = ref[rs:re], query[qs:qe]
        d = np.sqrt(((q[:, None, :] - rr[None, :, :]) ** 2).sum(-1)) if qe > qs and re > rs \
            else np.zeros((qe - qs, re - rs))
        for i in range(qe - qs):
            nz = np.nonzero(d[i] <= r)[0]
            idx_parts.append(nz + rs)
            d_parts.append(d[i, nz])
            counts.append(len(nz))
    splits = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=splits[1:])
    cat = (lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dt))
    return cat(idx_parts, np.int64), cat(d_parts, np.float64), splits
