# SPDX-License-Identifier: Apache-2.0
"""Strided reduction of a ``Voxels`` (drop-in for warpconvnet/nn/functional/sparse_pool.py:25-141;
what ``stride_mode=REDUCE_AND_STRIDE`` runs in front of the conv, helper.py:275-288).

The reference sorts the CSR kernel map by output row (``to_csr``) and calls
``torch_scatter.segment_csr`` on the gathered rows. The kernel map built here already carries the
dense ``pair_table[K, M]`` (input row of offset k for every output row, -1 when absent), so max /
min / sum / mean walk its K rows with O(M·C) temporaries: no sort, no ``[L, C]`` gather, and a
deterministic result. The other reductions go through ``to_csr`` + ``row_reduction`` like the
reference. Output rows without any input stay zero (sparse_pool.py:92-108). On exact ties the max /
min gradient is split between the tied inputs (``torch.maximum``); ``segment_csr`` sends it to one.
Device-side torch glue around the native kernel map — pooling itself is outside the hot path.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch
from torch import Tensor

from warpconvnet_b200.geometry.coords.integer import IntCoords
from warpconvnet_b200.geometry.coords.ops.stride import stride_coords
from warpconvnet_b200.geometry.coords.search.cache import IntSearchCache, IntSearchCacheKey
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
from warpconvnet_b200.geometry.types.voxels import Voxels
from warpconvnet_b200.ops.reductions import REDUCTIONS, row_reduction
from warpconvnet_b200.utils.ntuple import ntuple


def reduce_over_pair_table(features: Tensor, pair_table: Tensor, reduction: REDUCTIONS) -> Tensor:
    """``out[m] = reduce_k features[pair_table[k, m]]`` over the present (>= 0) entries."""
    K, M = pair_table.shape
    present = pair_table >= 0
    rows = pair_table.clamp_min(0).long()
    if reduction in (REDUCTIONS.MAX, REDUCTIONS.MIN):
        pick = torch.maximum if reduction == REDUCTIONS.MAX else torch.minimum
        fill = float("-inf") if reduction == REDUCTIONS.MAX else float("inf")
        acc = features.new_full((M, features.shape[1]), fill)
        for k in range(K):
            acc = pick(acc, features[rows[k]].masked_fill(~present[k, :, None], fill))
        return acc.masked_fill(~present.any(dim=0)[:, None], 0)
    acc = features.new_zeros((M, features.shape[1]), dtype=torch.float32)
    for k in range(K):
        acc = acc + features[rows[k]].float() * present[k, :, None]
    if reduction == REDUCTIONS.MEAN:
        acc = acc / present.sum(dim=0).clamp_min(1)[:, None]
    return acc.to(features.dtype)


def sparse_reduce(voxels: Voxels, kernel_size: Union[int, Tuple[int, ...]],
                  stride: Optional[Union[int, Tuple[int, ...]]] = None,
                  reduction: Union[REDUCTIONS, str] = REDUCTIONS.MAX, order=None) -> Voxels:
    """Pool ``voxels`` over ``kernel_size`` windows placed every ``stride`` voxels."""
    if isinstance(reduction, str):
        reduction = REDUCTIONS(reduction)
    nd = voxels.num_spatial_dims
    kernel_size = ntuple(kernel_size, ndim=nd)
    stride = kernel_size if stride is None else ntuple(stride, ndim=nd)
    in_stride = voxels.tensor_stride or ntuple(1, ndim=nd)
    out_stride = tuple(o * s for o, s in zip(stride, in_stride))

    bin_coords = voxels.batch_indexed_coordinates
    bout, out_offsets = stride_coords(bin_coords, stride, n_batches=len(voxels.offsets) - 1)
    key = IntSearchCacheKey(kernel_size=kernel_size, kernel_dilation=ntuple(1, ndim=nd),
                            transposed=False, generative=False,
                            stride_mode="STRIDED_CONV_MODE.STRIDE_ONLY",
                            skip_symmetric_kernel_map=False, in_offsets=voxels.offsets,
                            out_offsets=out_offsets)
    if not isinstance(voxels.cache, IntSearchCache):
        voxels._extra_attributes["_cache"] = IntSearchCache()
    kernel_map = voxels.cache.get(key)
    if kernel_map is None:
        kernel_map = generate_kernel_map(bin_coords, bout, stride, kernel_size,
                                         ntuple(1, ndim=nd))
        voxels.cache.put(key, kernel_map)

    feats = voxels.feature_tensor
    n_out = bout.shape[0]
    if reduction in (REDUCTIONS.MAX, REDUCTIONS.MIN, REDUCTIONS.SUM, REDUCTIONS.MEAN):
        out = reduce_over_pair_table(feats, kernel_map.pair_table(n_out), reduction)
    else:
        in_rows, out_rows, row_offsets = kernel_map.to_csr()
        out = feats.new_zeros((n_out, feats.shape[1]))
        out[out_rows] = row_reduction(feats[in_rows.long()], row_offsets, reduction)
    coords = IntCoords(bout[:, 1:].contiguous(), offsets=out_offsets.cpu())
    coords._bcoords = bout
    return voxels.replace(batched_coordinates=coords, batched_features=out,
                          tensor_stride=out_stride)


def sparse_max_pool(voxels: Voxels, kernel_size, stride=None) -> Voxels:
    return sparse_reduce(voxels, kernel_size, stride, reduction=REDUCTIONS.MAX)


def sparse_avg_pool(voxels: Voxels, kernel_size, stride=None) -> Voxels:
    return sparse_reduce(voxels, kernel_size, stride, reduction=REDUCTIONS.MEAN)
