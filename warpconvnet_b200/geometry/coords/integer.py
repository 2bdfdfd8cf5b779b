# SPDX-License-Identifier: Apache-2.0
"""Integer (voxel) coordinates (warpconvnet/geometry/coords/integer.py:23-211)."""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor

from warpconvnet_b200.geometry.base.batched import Coords, list_to_cat_tensor
from warpconvnet_b200.geometry.coords.ops.batch_index import (batch_indexed_coordinates,
                                                             offsets_from_batch_index)
from warpconvnet_b200.utils.ntuple import ntuple


class IntCoords(Coords):
    def __init__(self, batched_tensor, offsets=None, voxel_size: Optional[float] = None,
                 tensor_stride: Optional[Union[int, Tuple[int, ...]]] = None,
                 device: Optional[str] = None):
        if isinstance(batched_tensor, (list, tuple)):
            assert offsets is None, "If batched_tensors is a list, offsets must be None."
            batched_tensor, offsets, _ = list_to_cat_tensor(batched_tensor)
        if isinstance(offsets, (list, tuple)):
            offsets = torch.LongTensor(list(offsets))
        if offsets is None:
            offsets = torch.LongTensor([0, batched_tensor.shape[0]])
        if device is not None:
            batched_tensor = batched_tensor.to(device)
        self.offsets = offsets.cpu()
        self.batched_tensor = batched_tensor
        self.voxel_size = voxel_size
        self.tensor_stride = (ntuple(tensor_stride, ndim=self.batched_tensor.shape[1])
                              if tensor_stride is not None else None)
        self._bcoords: Optional[Tensor] = None
        self._hashmap = None
        self.check()

    def check(self):
        Coords.check(self)
        assert self.batched_tensor.dtype in (torch.int32, torch.int64), \
            "Discrete coordinates must be integers"

    def to(self, device=None, dtype=None):
        t = self.batched_tensor if device is None else self.batched_tensor.to(device)
        return self.__class__(t, self.offsets, voxel_size=self.voxel_size,
                              tensor_stride=self.tensor_stride)

    @property
    def batch_indexed_coordinates(self) -> Tensor:
        """[N, D+1] int32 (batch, coords...). Materialised once and cached — the reference
        rebuilds it with repeat_interleave + cat on every call (ops/batch_index.py:90-96)."""
        if self._bcoords is None or self._bcoords.device != self.batched_tensor.device:
            self._bcoords = batch_indexed_coordinates(self.batched_tensor.int(),
                                                      self.offsets).contiguous()
        return self._bcoords

    def unique(self) -> "IntCoords":
        from warpconvnet_b200.geometry.coords.ops.stride import unique_with_offsets
        uniq, _, offs = unique_with_offsets(self.batch_indexed_coordinates, len(self.offsets) - 1)
        return self.__class__(uniq[:, 1:].contiguous(), offs,
                              voxel_size=self.voxel_size, tensor_stride=self.tensor_stride)

    def prune(self, mask: Tensor) -> "IntCoords":
        """Rows where ``mask`` is true, offsets recomputed per batch item (integer.py:99-134)."""
        assert mask.shape[0] == self.batched_tensor.shape[0], "Mask must match tensor shape"
        mask = mask.to(self.batched_tensor.device)
        if mask.dtype != torch.bool:
            mask = mask.bool()
        from warpconvnet_b200.geometry.coords.ops.batch_index import batch_index_from_offset
        bidx = batch_index_from_offset(self.offsets, device=self.batched_tensor.device)
        offsets = offsets_from_batch_index(bidx[mask], len(self.offsets) - 1)
        return self.__class__(self.batched_tensor[mask], offsets.to(self.offsets.dtype),
                              voxel_size=self.voxel_size, tensor_stride=self.tensor_stride)

    def sort(self, ordering=None) -> "IntCoords":
        """Rows of every batch item in z-order (integer.py ``sort``)."""
        from warpconvnet_b200.geometry.coords.ops.serialization import POINT_ORDERING, encode
        res = encode(self.batched_tensor, batch_offsets=self.offsets,
                     order=POINT_ORDERING.MORTON_XYZ if ordering is None else ordering,
                     return_perm=True)
        return self.__class__(self.batched_tensor[res.perm], self.offsets,
                              voxel_size=self.voxel_size, tensor_stride=self.tensor_stride)

    def expand(self, kernel_size, dilation=1) -> "IntCoords":
        from warpconvnet_b200.geometry.coords.ops.expand import expand_coords
        nd = self.num_spatial_dims
        out, offs = expand_coords(self.batch_indexed_coordinates, ntuple(kernel_size, ndim=nd),
                                  ntuple(dilation, ndim=nd), n_batches=len(self.offsets) - 1)
        return self.__class__(out[:, 1:].contiguous(), offs.to(self.offsets.dtype),
                              voxel_size=self.voxel_size, tensor_stride=self.tensor_stride)

    @property
    def hashmap(self):
        from warpconvnet_b200.geometry.coords.search.packed_hashmap import PackedHashTable
        if self._hashmap is None:
            bc = self.batch_indexed_coordinates
            if bc.shape[1] == 3:
                bc = torch.nn.functional.pad(bc, (0, 1), value=0)
            self._hashmap = PackedHashTable.from_coords(bc)
        return self._hashmap

    @property
    def stride(self):
        return self.tensor_stride

    def set_tensor_stride(self, tensor_stride):
        self.tensor_stride = ntuple(tensor_stride, ndim=self.num_spatial_dims)


class RealCoords(Coords):
    """Continuous point coordinates (warpconvnet/geometry/coords/real.py)."""

    def __init__(self, batched_tensor, offsets=None, device: Optional[str] = None):
        super().__init__(batched_tensor, offsets, device=device)
