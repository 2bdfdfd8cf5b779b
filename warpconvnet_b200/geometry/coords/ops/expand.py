# SPDX-License-Identifier: Apache-2.0
"""Generative coordinate expansion (warpconvnet/geometry/coords/ops/expand.py:17-75): the union of
``coord + offset_k`` over all kernel offsets, deduplicated, sorted by (batch, x, y, z). CUDA
tensors run the native chain of ``csrc/coords.cu`` (keys for all K x N candidates -> sort ->
compaction) without a host sync on the compute stream; the reference loops over kernel batches in
Python with a hash insert + ``.cpu()`` per batch."""
from typing import Optional, Tuple

import torch
from torch import Tensor

from warpconvnet_b200.geometry.coords.ops.batch_index import offsets_from_batch_index
from warpconvnet_b200.geometry.coords.ops.stride import (_check_range_cpu, _unique_rows_cpu,
                                                         unique_rows_device)


@torch.no_grad()
def expand_coords(batch_indexed_coords: Tensor, kernel_size: Tuple[int, ...],
                  kernel_dilation: Tuple[int, ...], n_batches: Optional[int] = None
                  ) -> Tuple[Tensor, Tensor]:
    from warpconvnet_b200.geometry.coords.search.torch_discrete import kernel_offsets_from_size
    offs = kernel_offsets_from_size(kernel_size, kernel_dilation,
                                    device=batch_indexed_coords.device)
    pad = batch_indexed_coords.shape[1] == 3
    if batch_indexed_coords.is_cuda:
        c = torch.nn.functional.pad(batch_indexed_coords, (0, 1), value=0) if pad \
            else batch_indexed_coords
        o3 = offs[:, 1:]
        if pad:
            o3 = torch.nn.functional.pad(o3, (0, 1), value=0)
        nb = 512 if n_batches is None else n_batches
        uniq, offsets, _ = unique_rows_device(c, (1, 1, 1), o3.contiguous().int(), nb)
        if n_batches is None:
            last = int((offsets[1:] - offsets[:-1]).nonzero().max()) + 1 if uniq.shape[0] else 0
            offsets = offsets[:last + 1]
        return (uniq[:, :3] if pad else uniq).contiguous(), offsets
    allc = (batch_indexed_coords[None, :, :] + offs[:, None, :]).reshape(-1, offs.shape[1]).int()
    c4 = torch.nn.functional.pad(allc, (0, 1), value=0) if pad else allc
    _check_range_cpu(c4)
    uniq, _ = _unique_rows_cpu(c4)
    if pad:
        uniq = uniq[:, :3]
    nb = n_batches
    if nb is None:
        nb = int(batch_indexed_coords[:, 0].max()) + 1 if batch_indexed_coords.numel() else 0
    return uniq.contiguous(), offsets_from_batch_index(uniq[:, 0], nb)
