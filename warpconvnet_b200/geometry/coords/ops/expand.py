# SPDX-License-Identifier: Apache-2.0
"""Generative coordinate expansion (warpconvnet/geometry/coords/ops/expand.py:17-75): the union of
``coord + offset_k`` over all kernel offsets, deduplicated, batch-sorted."""
from typing import Tuple

import torch
from torch import Tensor

from warpconvnet_b200.geometry.coords.ops.batch_index import offsets_from_batch_index
from warpconvnet_b200.geometry.coords.ops.stride import unique_coords


@torch.no_grad()
def expand_coords(batch_indexed_coords: Tensor, kernel_size: Tuple[int, ...],
                  kernel_dilation: Tuple[int, ...]) -> Tuple[Tensor, Tensor]:
    from warpconvnet_b200.geometry.coords.search.torch_discrete import kernel_offsets_from_size
    offs = kernel_offsets_from_size(kernel_size, kernel_dilation,
                                    device=batch_indexed_coords.device)
    allc = (batch_indexed_coords[None, :, :] + offs[:, None, :]).reshape(-1, offs.shape[1])
    uniq, _ = unique_coords(allc.int())
    nb = int(batch_indexed_coords[:, 0].max().item()) + 1 if batch_indexed_coords.numel() else 0
    return uniq.contiguous(), offsets_from_batch_index(uniq[:, 0], nb)
