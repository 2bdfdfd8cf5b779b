# SPDX-License-Identifier: Apache-2.0
"""Batch-index helpers (warpconvnet/geometry/coords/ops/batch_index.py:18-31,90-130)."""
import torch
from torch import Tensor


@torch.no_grad()
def batch_index_from_offset(offsets: Tensor, device=None) -> Tensor:
    """int64 batch index per row. Built ON the target device from the (CPU) offsets: a few fill
    launches, no host-side repeat_interleave and no pageable H2D copy of an N-element tensor
    (the reference does both, ops/batch_index.py:18-31)."""
    assert len(offsets) > 1, "offsets must have at least two elements. [0, N] for batch size 1"
    offs = [int(v) for v in offsets.tolist()]
    dev = torch.device(device) if device is not None else torch.device("cpu")
    if dev.type == "cpu":
        count = torch.diff(offsets.cpu())
        return torch.arange(len(count), dtype=torch.long).repeat_interleave(count)
    out = torch.empty(offs[-1], dtype=torch.long, device=dev)
    for b in range(len(offs) - 1):
        out[offs[b]:offs[b + 1]] = b
    return out


@torch.no_grad()
def batch_indexed_coordinates(batched_coords: Tensor, offsets: Tensor) -> Tensor:
    """[N, D+1] (batch, coords...) in the dtype of the coordinates."""
    if batched_coords.device.type == "cpu":
        batch_index = batch_index_from_offset(offsets).to(batched_coords.dtype)
        return torch.cat([batch_index.unsqueeze(1), batched_coords], dim=1)
    n, d = batched_coords.shape
    offs = [int(v) for v in offsets.tolist()]
    out = torch.empty((n, d + 1), dtype=batched_coords.dtype, device=batched_coords.device)
    out[:, 1:] = batched_coords
    for b in range(len(offs) - 1):
        out[offs[b]:offs[b + 1], 0] = b
    return out


@torch.no_grad()
def offsets_from_batch_index(batch_index: Tensor, batch_size=None) -> Tensor:
    """CPU int64 offsets[B+1] from a sorted batch-index column."""
    if batch_index.numel() == 0:
        return torch.zeros(1 if batch_size is None else batch_size + 1, dtype=torch.int64)
    counts = torch.bincount(batch_index.long(), minlength=0 if batch_size is None else batch_size)
    counts = counts.cpu()
    return torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(counts, 0)])
