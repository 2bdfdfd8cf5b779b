# SPDX-License-Identifier: Apache-2.0
"""Batch-index helpers (warpconvnet/geometry/coords/ops/batch_index.py:18-31,90-130)."""
import torch
from torch import Tensor


@torch.no_grad()
def batch_index_from_offset(offsets: Tensor, device=None) -> Tensor:
    assert len(offsets) > 1, "offsets must have at least two elements. [0, N] for batch size 1"
    count = torch.diff(offsets.cpu())
    batch = torch.arange(len(count), dtype=torch.long).repeat_interleave(count)
    return batch.to(device) if device is not None else batch


@torch.no_grad()
def batch_indexed_coordinates(batched_coords: Tensor, offsets: Tensor) -> Tensor:
    batch_index = batch_index_from_offset(offsets, device=batched_coords.device).to(batched_coords.dtype)
    return torch.cat([batch_index.unsqueeze(1), batched_coords], dim=1)


@torch.no_grad()
def offsets_from_batch_index(batch_index: Tensor, batch_size=None) -> Tensor:
    """CPU int64 offsets[B+1] from a sorted batch-index column."""
    if batch_index.numel() == 0:
        return torch.zeros(1 if batch_size is None else batch_size + 1, dtype=torch.int64)
    counts = torch.bincount(batch_index.long(), minlength=0 if batch_size is None else batch_size)
    counts = counts.cpu()
    return torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(counts, 0)])
