# SPDX-License-Identifier: Apache-2.0
"""Strided (downsampled) output coordinates and coordinate-set uniqueness
(drop-in for warpconvnet/geometry/coords/ops/stride.py:18-56 and the ``unique`` used by
geometry/types/voxels.py:271-278).

The reference floor-divides through float, dedups with a racy hash insert and then argsorts the
batch column (row order is nondeterministic), with several host syncs. Here CUDA tensors go through
ONE native chain (``csrc/coords.cu``: key generation -> radix sort -> head flags -> scan ->
compaction + per-batch offsets) that never synchronises the compute stream; rows come out sorted
by (batch, x, y, z) — deterministic, already batch-contiguous, spatially coherent for the next
level's gathers. CPU tensors (host-side tests, data preparation) use the same rule in torch.

The only thing the host has to learn is the row count. In eager mode the chain runs on a side
stream and the host waits for ITS event only (the compute stream keeps its backlog, so the host
keeps running ahead of the GPU through a MinkUNet level change); under CUDA-graph capture the
count comes from the ``SizeTape`` recorded by an eager warm-up pass (``utils/graph.py``).
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from warpconvnet_b200.utils.graph import active_tape
from warpconvnet_b200.utils.ntuple import ntuple

_OFF = 1 << 17  # bias that maps [-131072, 131071] to [0, 262143]
_COORD_MIN, _COORD_MAX, _BATCH_MAX = -131072, 131071, 511


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


def _range_error() -> ValueError:
    return ValueError(f"Coordinate out of range: batch must be in [0, {_BATCH_MAX}] and spatial "
                      f"coords in [{_COORD_MIN}, {_COORD_MAX}]")


_SIDE = {}


def _side_stream(dev: torch.device) -> "torch.cuda.Stream":
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _SIDE:
        _SIDE[idx] = torch.cuda.Stream(device=idx)
    return _SIDE[idx]


@torch.no_grad()
def unique_rows_device(bcoords: Tensor, stride: Tuple[int, int, int] = (1, 1, 1),
                       offsets3: Optional[Tensor] = None, n_batches: Optional[int] = None,
                       want_index: bool = False):
    """(rows [M, 4] int32 sorted by (b, x, y, z), CPU int64 offsets [B + 1], first source row of
    every output row | None) of unique{(b, floor(xyz / stride) + offset_k)} on a CUDA tensor."""
    from warpconvnet_b200 import _ops
    assert bcoords.is_cuda and bcoords.shape[1] == 4
    bc = bcoords if (bcoords.dtype == torch.int32 and bcoords.is_contiguous()) \
        else bcoords.int().contiguous()
    n = bc.shape[0]
    K = 1 if offsets3 is None else offsets3.shape[0]
    if n_batches is None:
        n_batches = _BATCH_MAX + 1  # unknown: every possible batch item gets an offset entry
    signature = ("coords_unique", n, tuple(int(s) for s in stride), K, n_batches, bool(want_index))
    tape = active_tape()
    dev = bc.device
    main = torch.cuda.current_stream(dev)
    if torch.cuda.is_current_stream_capturing():
        if tape is None or not tape.replaying:
            raise RuntimeError(
                "creating a coordinate set (strided / generative conv, unique) under CUDA-graph "
                "capture needs the sizes of an eager warm-up pass: wrap the capture in "
                "warpconvnet_b200.utils.graph.SizeTape.replay() (see capture_step)")
        total, offs_list, status = tape.next(signature)
        rows, first, meta = _ops.coords_unique(bc, stride, offsets3, n_batches, want_index)
        tape.expect(meta[n_batches + 1], total)
    else:
        side = _side_stream(dev)
        ready = getattr(bcoords, "_wcn_ready", None)
        if ready is None:
            # produced on the compute stream (user input): order the side stream behind it
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
        else:
            side.wait_event(ready)  # produced by an earlier chain on the side stream itself
        with torch.cuda.stream(side):
            rows, first, meta = _ops.coords_unique(bc, stride, offsets3, n_batches, want_index)
            host = torch.empty(n_batches + 3, dtype=torch.int32, pin_memory=True)
            host.copy_(meta, non_blocking=True)
            done = torch.cuda.Event()
            done.record(side)
        bc.record_stream(side)
        done.synchronize()  # waits for the side stream's chain only
        vals = host.tolist()
        total, status = vals[n_batches + 1], vals[n_batches + 2]
        offs_list = vals[:n_batches + 1]
        if status & 2:
            raise _range_error()
        if tape is not None and tape.recording:
            tape.append(signature, (total, offs_list, status))
    if total * 2 < rows.shape[0]:
        # release the upper-bound buffer (generative expansion: K x larger than the result)
        with torch.cuda.stream(main if torch.cuda.is_current_stream_capturing() else side):
            rows = rows[:total].clone()
            first = first[:total].clone() if first is not None else None
    else:
        rows = rows[:total]
        first = first[:total] if first is not None else None
    if not torch.cuda.is_current_stream_capturing():
        fin = torch.cuda.Event()
        fin.record(side)
        main.wait_event(fin)           # consumers on the compute stream are ordered behind it
        for t in (rows, first):
            if t is not None:
                t.record_stream(main)
        rows._wcn_ready = fin          # lets the next level's chain skip the compute stream
    return rows, torch.tensor(offs_list, dtype=torch.int64), first


@torch.no_grad()
def _unique_rows_cpu(bcoords: Tensor):
    keys = pack_sortable(bcoords)
    order = torch.argsort(keys, stable=True)
    sk = keys[order]
    first = torch.ones_like(sk, dtype=torch.bool)
    first[1:] = sk[1:] != sk[:-1]
    return unpack_sortable(sk[first]), order[first]


def _check_range_cpu(c: Tensor) -> None:
    if c.numel() and (int(c[:, 0].min()) < 0 or int(c[:, 0].max()) > _BATCH_MAX
                      or int(c[:, 1:].min()) < _COORD_MIN or int(c[:, 1:].max()) > _COORD_MAX):
        raise _range_error()


@torch.no_grad()
def unique_with_offsets(bcoords: Tensor, n_batches: int) -> Tuple[Tensor, Tensor, Tensor]:
    """(unique rows sorted by key, index of the first occurrence of each, CPU offsets [B + 1])."""
    pad = bcoords.shape[1] == 3
    c = torch.nn.functional.pad(bcoords, (0, 1), value=0) if pad else bcoords
    if c.is_cuda:
        out, offs, idx = unique_rows_device(c, n_batches=n_batches, want_index=True)
        idx = idx.long()
    else:
        _check_range_cpu(c)
        out, idx = _unique_rows_cpu(c)
        offs = _offsets_from_sorted_batch(out[:, 0], n_batches)
    return (out[:, :3] if pad else out), idx, offs


@torch.no_grad()
def unique_coords(bcoords: Tensor, n_batches: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    """(unique rows sorted by key, index of the first occurrence of each)."""
    out, idx, _ = unique_with_offsets(bcoords, _BATCH_MAX + 1 if n_batches is None else n_batches)
    return out, idx


def _offsets_from_sorted_batch(batch: Tensor, n_batches: Optional[int]) -> Tensor:
    from warpconvnet_b200.geometry.coords.ops.batch_index import offsets_from_batch_index
    return offsets_from_batch_index(batch, n_batches)


@torch.no_grad()
def stride_coords(batch_indexed_coords: Tensor, stride: Tuple[int, ...], order=None,
                  n_batches: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    """(unique floor(coords / stride) rows sorted by (batch, x, y, z), CPU offsets [B + 1]).
    ``n_batches``: number of batch items when the caller knows it (offsets then always have
    B + 1 entries, also for trailing empty items)."""
    num_spatial_dims = batch_indexed_coords.shape[1] - 1
    stride = ntuple(stride, ndim=num_spatial_dims)
    if all(s == 1 for s in stride):
        return batch_indexed_coords, _offsets_from_sorted_batch(batch_indexed_coords[:, 0],
                                                                n_batches)
    pad = num_spatial_dims == 2
    c = torch.nn.functional.pad(batch_indexed_coords, (0, 1), value=0) if pad \
        else batch_indexed_coords
    if c.is_cuda:
        nb = n_batches
        if nb is None:
            # number of batch items = last batch index + 1: the rows are batch-sorted, and the
            # caller (spatially_sparse_conv) normally passes it; fall back to the full range
            nb = _BATCH_MAX + 1
        out, offs, _ = unique_rows_device(c, tuple(stride) + ((1,) if pad else ()), None, nb)
        if n_batches is None:
            # trim the trailing empty batch items of the full-range layout
            last = int((offs[1:] - offs[:-1]).nonzero().max()) + 1 if out.shape[0] else 0
            offs = offs[:last + 1]
    else:
        _check_range_cpu(c)
        q = c.clone()
        for d, s in enumerate(stride):
            q[:, d + 1] = torch.div(c[:, d + 1], int(s), rounding_mode="floor")
        out, _ = _unique_rows_cpu(q)
        offs = _offsets_from_sorted_batch(out[:, 0], n_batches)
    if pad:
        out = out[:, :3]
    return out.contiguous(), offs
