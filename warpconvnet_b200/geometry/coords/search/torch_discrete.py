# SPDX-License-Identifier: Apache-2.0
"""Kernel-map generation for sparse convolution (drop-in for
warpconvnet/geometry/coords/search/torch_discrete.py:23-57,296-432).

Pipeline (all on the current CUDA stream, NO host synchronisation on the conv path):
  hash build -> K-offset probe (pair table + per-block counts + per-row offset mask)
  -> block scan / offsets -> async D2H of (offsets, status) -> deterministic CSR scatter.
The host only waits when `offsets` / `in_maps` / `out_maps` are read on the CPU
(IntSearchResult._resolve). The reference needs >= 6 host syncs for the same work
(SURVEY.md §3.1).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from warpconvnet_b200 import _ops
from warpconvnet_b200.utils.ntuple import ntuple
from .packed_hashmap import PackedHashTable
from .search_results import HostCopy, IntSearchResult, check_pending_kernel_maps

_OFFSET_CACHE: Dict[tuple, Tensor] = {}
_STATS_ON_SIDE = False  # True: statistics pass of a submanifold map on the CSR side stream (no gain
# in graph replays - the sort then derives the masks itself - and noisy eager / e2e steps: off)
# Upper-bound CSR buffers of at most 2 x 512 MiB (K * M int32 each; the real pair count is ~1/3 of
# it on surface data): a 27-offset map of 2.4 M voxels (MinkUNet-14 full resolution, 8 scenes)
# stays on the sync-free path. Above it the exact length is read back (one host sync).
_DEFERRED_MAX_PAIRS = 1 << 27


_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def _side_stream(dev: torch.device) -> "torch.cuda.Stream":
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    s = _SIDE_STREAMS.get(idx)
    if s is None:
        s = torch.cuda.Stream(device=idx)
        _SIDE_STREAMS[idx] = s
    return s


@torch.no_grad()
def kernel_offsets_from_size(kernel_size: Tuple[int, ...], kernel_dilation: Tuple[int, ...],
                             center_offset: Optional[Tuple[int, ...]] = None,
                             device: Optional[torch.device] = None) -> Tensor:
    """[K, D+1] int32 offsets, batch column first (torch_discrete.py:23-57): 'ij' meshgrid order,
    odd sizes are centred, even sizes start at 0."""
    assert len(kernel_size) == len(kernel_dilation)
    ranges = [torch.arange(int(s), dtype=torch.int32) for s in kernel_size]
    grids = torch.meshgrid(*ranges, indexing="ij")
    if center_offset is None:
        center_offset = [(s - 1) // 2 if s % 2 == 1 else 0 for s in kernel_size]
    assert len(center_offset) == len(kernel_size)
    cols = [(g.flatten() - int(center_offset[i])) * int(kernel_dilation[i])
            for i, g in enumerate(grids)]
    cols = [torch.zeros_like(cols[0])] + cols
    return torch.stack(cols, dim=1).contiguous().to(device)


def _offsets3(kernel_size, kernel_dilation, center_offset, device) -> Tensor:
    key = (tuple(kernel_size), tuple(kernel_dilation),
           None if center_offset is None else tuple(center_offset), str(device))
    t = _OFFSET_CACHE.get(key)
    if t is None:
        t = kernel_offsets_from_size(kernel_size, kernel_dilation, center_offset)[:, 1:]
        t = t.contiguous().to(device)
        _OFFSET_CACHE[key] = t
    return t


@torch.no_grad()
def generate_kernel_map(
    batch_indexed_in_coords: Tensor,
    batch_indexed_out_coords: Tensor,
    in_to_out_stride_ratio: Tuple[int, ...],
    kernel_size: Tuple[int, ...],
    kernel_dilation: Optional[Tuple[int, ...]] = None,
    kernel_center_offset: Optional[Tuple[int, ...]] = None,
    method: str = "size",
    skip_symmetric_kernel_map: bool = False,
    same_coords: Optional[bool] = None,
    build_plan: bool = True,
    **kwargs,
) -> IntSearchResult:
    """``coord(in) = stride * coord(out) + offset[k]`` pairs for every kernel offset k.

    Returns an ``IntSearchResult`` whose CSR rows are in ascending output-row order inside each
    offset (the reference's order is unspecified). ``same_coords=True`` tells the builder that the
    two coordinate tensors are the same set in the same order (submanifold conv) so the dgrad
    pass can reuse the forward tables; when None it is inferred from tensor identity.
    ``build_plan`` also builds the forward tile plan here (overlapped with the CSR emission).
    """
    assert batch_indexed_in_coords.dtype == torch.int32
    assert batch_indexed_out_coords.dtype == torch.int32
    dev = batch_indexed_in_coords.device
    assert dev == batch_indexed_out_coords.device
    if skip_symmetric_kernel_map:
        raise NotImplementedError(
            "skip_symmetric_kernel_map is not supported (the reference's conv path never sets it, "
            "helper.py:446-455)")
    if same_coords is None:
        same_coords = batch_indexed_in_coords is batch_indexed_out_coords or (
            batch_indexed_in_coords.data_ptr() == batch_indexed_out_coords.data_ptr()
            and batch_indexed_in_coords.shape == batch_indexed_out_coords.shape)

    if not torch.cuda.is_current_stream_capturing():
        check_pending_kernel_maps()  # non-blocking: raises deferred errors of earlier maps

    in_c, out_c = batch_indexed_in_coords, batch_indexed_out_coords
    kernel_size = tuple(int(k) for k in kernel_size)
    stride = tuple(int(s) for s in in_to_out_stride_ratio)
    if in_c.shape[1] == 3:  # 2-D conv: pad z = 0 (torch_discrete.py:328-342)
        in_c = torch.nn.functional.pad(in_c, (0, 1), value=0)
        out_c = in_c if same_coords else torch.nn.functional.pad(out_c, (0, 1), value=0)
        kernel_size = kernel_size + (1,)
        stride = stride + (1,)
        if kernel_dilation is not None:
            kernel_dilation = tuple(kernel_dilation) + (1,)
        if kernel_center_offset is not None:
            kernel_center_offset = tuple(kernel_center_offset) + (0,)
    assert in_c.shape[1] == 4, f"Expected 4D batch-indexed coords, got {in_c.shape[1]}D"
    assert len(stride) == 3
    if kernel_dilation is None:
        kernel_dilation = (1, 1, 1)
    kernel_dilation = ntuple(kernel_dilation, ndim=3)

    in_c = in_c.contiguous()
    out_c = out_c.contiguous()
    n_in, n_out = in_c.shape[0], out_c.shape[0]
    K = int(np.prod(kernel_size))

    identity_map_index = None
    is_odd = all(k % 2 == 1 for k in kernel_size)
    if is_odd and n_in == n_out:  # by COUNT, exactly like torch_discrete.py:363-370
        identity_map_index = K // 2

    table = PackedHashTable.from_coords(in_c, check=False)
    offs3 = _offsets3(kernel_size, kernel_dilation, kernel_center_offset, dev)
    symmetric = bool(same_coords and is_odd and all(s == 1 for s in stride)
                     and kernel_center_offset is None)
    stats_on_side = False
    if symmetric and n_out > 0:
        # Table only on the compute stream: the tile plan derives the row masks inside its sort
        # kernel, and the per-block pair counts are only needed by the CSR branch, which runs on
        # the side stream — the statistics pass over the table leaves the critical path.
        stats_on_side = _STATS_ON_SIDE and build_plan and K <= 32 and n_out <= (1 << 20)
        pair_table, block_counts, mask_keys = _ops.kernel_map_search_symmetric(
            table.keys_tensor, table.values_tensor, out_c, offs3, table.status_tensor,
            with_stats=not stats_on_side)
    else:
        pair_table, block_counts, mask_keys = _ops.kernel_map_search(
            table.keys_tensor, table.values_tensor, out_c, offs3, stride)
    # Fork: the CSR branch (block scan -> offsets -> async D2H of offsets/status -> deterministic
    # scatter) runs on a side stream while the main stream builds the forward tile plan (radix
    # sort of the row masks -> step lists); both only read the pair table. Joined before returning,
    # so callers (and CUDA-graph capture) see one stream.
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    deferred = K * n_out <= _DEFERRED_MAX_PAIRS
    fork = torch.cuda.Event()
    fork.record(main)
    side.wait_event(fork)
    with torch.cuda.stream(side):
        if stats_on_side:
            block_counts, _ = _ops.kernel_map_stats(pair_table, want_mask=False)
        offsets_dev = _ops.kernel_map_count(block_counts)
        if deferred:
            # No host sync: the CSR lists go into upper-bound sized buffers, (offsets, status)
            # travel to pinned host memory asynchronously and are only waited for when somebody
            # reads `offsets` / `in_maps` / `out_maps` on the host (IntSearchResult._resolve).
            meta = torch.cat([offsets_dev, table.status_tensor])
            # under capture: read back (synchronously) only if somebody asks later
            host = HostCopy(meta, asynchronous=not torch.cuda.is_current_stream_capturing())
            in_maps, out_maps = _ops.kernel_map_scatter(pair_table, block_counts, offsets_dev,
                                                        K * n_out)
            result = IntSearchResult._from_device(in_maps, out_maps, offsets_dev, host,
                                                  table.raise_if_failed, identity_map_index)
            host = meta
        else:
            # very large K * M: allocate the exact length instead (one host sync)
            host = torch.cat([offsets_dev, table.status_tensor]).cpu()
            table.raise_if_failed(int(host[-1]))
            offsets_cpu = host[:-1].clone()
            num_pairs = int(offsets_cpu[-1])
            in_maps, out_maps = _ops.kernel_map_scatter(pair_table, block_counts, offsets_dev,
                                                        num_pairs)
            result = IntSearchResult(in_maps, out_maps, offsets_cpu,
                                     identity_map_index=identity_map_index)
        join = torch.cuda.Event()
        join.record(side)
    for t in (pair_table, table.status_tensor) + (() if stats_on_side else (block_counts,)):
        t.record_stream(side)       # allocated on the main stream, read by the side stream
    for t in (offsets_dev, in_maps, out_maps) + ((host,) if host.is_cuda else ()) + (
            (block_counts,) if stats_on_side else ()):
        t.record_stream(main)       # allocated on the side stream, consumed on the main stream
    if build_plan and n_out > 0:
        result._fwd_plan = _ops.build_tile_plan(pair_table, mask_keys)
        mask_keys = None
    main.wait_event(join)
    result._offsets_dev = offsets_dev
    result._block_prefix = block_counts  # scanned in place: pairs of offset k with out row < 256*b
    result._pair_table = pair_table
    result._mask_keys = mask_keys
    result._n_in, result._n_out = n_in, n_out
    result._symmetric = bool(same_coords and is_odd and all(s == 1 for s in stride)
                             and all(d == 1 for d in kernel_dilation)
                             and kernel_center_offset is None)
    result._hashtable = table
    result._kernel_size = kernel_size
    return result


def _int_sequence_hash(arr: Sequence[int]) -> int:
    x = hash(arr[0])
    for i in range(1, len(arr)):
        x = (x * 31 + hash(arr[i])) & 0xFFFFFFFF
    return x
