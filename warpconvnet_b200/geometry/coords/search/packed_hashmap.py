# SPDX-License-Identifier: Apache-2.0
"""Packed 4-D coordinate hash table on the GPU.

Same user-visible contract as the reference's ``PackedHashTable``
(warpconvnet/geometry/coords/search/packed_hashmap.py:27-260, ``_packed_base.py:32-134``):
``from_coords`` / ``insert`` / ``search`` / ``keys_tensor`` / ``values_tensor`` / ``capacity``,
key layout 1|9|18|18|18 bits, capacity = next_pow2(max(16, 2N)), value = insertion index.

B200-native differences: the coordinate range check runs inside the insert kernel and is
reported through the status word together with "table full", so building a table needs no host
synchronisation at all (the reference does 4 ``.item()`` range checks + 1 status read,
packed_hashmap.py:72-81, _packed_base.py:115). ``status`` is checked lazily by
``raise_if_failed`` — ``generate_kernel_map`` folds it into the one D2H copy it needs anyway.
"""
from __future__ import annotations

import enum
from typing import Optional, Union

import torch
from torch import Tensor

from warpconvnet_b200 import _ops


class SearchMode(enum.IntEnum):
    LINEAR = 0
    DOUBLE_HASH = 1  # accepted for API compatibility; the table is always linear-probed
    WARP_COOP = 2


class PackedHashTable:
    BATCH_MAX = 511
    COORD_MIN = -131072
    COORD_MAX = 131071

    def __init__(self, capacity: int, device: Union[str, torch.device] = "cuda",
                 use_double_hash: bool = False):
        self._capacity = _ops.next_power_of_2(capacity)
        self._device = torch.device(device)
        self._keys: Optional[Tensor] = None
        self._values: Optional[Tensor] = None
        self._coords: Optional[Tensor] = None
        self._status: Optional[Tensor] = None
        self._num_entries = 0

    # ---- accessors (same names as the reference) ------------------------------------------
    @property
    def capacity(self) -> int:
        return self._capacity

    @property
    def device(self) -> torch.device:
        return self._keys.device if self._keys is not None else self._device

    @property
    def num_entries(self) -> int:
        return self._num_entries

    @property
    def keys_tensor(self) -> Tensor:
        return self._keys

    @property
    def values_tensor(self) -> Tensor:
        return self._values

    @property
    def status_tensor(self) -> Tensor:
        return self._status

    @property
    def key_dim(self) -> int:
        return 4

    @property
    def vector_keys(self) -> Tensor:
        if self._coords is None:
            raise RuntimeError("No coordinates stored. Call insert() first.")
        return self._coords[: self._num_entries]

    # ---- build / query ---------------------------------------------------------------------
    def insert(self, coords: Tensor, check: bool = True) -> None:
        assert coords.is_cuda, "coords must be on CUDA"
        assert coords.ndim == 2 and coords.shape[1] == 4
        coords = coords.contiguous().to(dtype=torch.int32, device=self._device)
        n = coords.shape[0]
        assert n <= self._capacity // 2, f"num_keys={n} exceeds capacity/2={self._capacity // 2}"
        self._keys = torch.empty(self._capacity, dtype=torch.int64, device=self._device)
        self._values = torch.empty(self._capacity, dtype=torch.int32, device=self._device)
        self._status = torch.zeros(1, dtype=torch.int32, device=self._device)
        _ops.hash_prepare(self._keys, self._values)
        _ops.hash_insert(self._keys, self._values, coords, self._status)
        self._num_entries = n
        self._coords = coords
        if check:
            self.raise_if_failed(int(self._status.item()))

    def raise_if_failed(self, status: int) -> None:
        """Same exceptions as the reference: ValueError for out-of-range coordinates
        (packed_hashmap.py:72-81), RuntimeError when the table is full (_packed_base.py:115-120)."""
        if status & 2:
            raise ValueError(
                f"Coordinate out of range: batch must be in [0, {self.BATCH_MAX}] and spatial "
                f"coords in [{self.COORD_MIN}, {self.COORD_MAX}]")
        if status & 1:
            raise RuntimeError(
                f"PackedHashTable.insert failed: hash table is full (num_keys={self._num_entries}, "
                f"capacity={self._capacity}). Increase capacity or reduce load factor.")

    @classmethod
    def from_coords(cls, coords: Tensor, device: Union[str, torch.device, None] = None,
                    capacity: Optional[int] = None, use_double_hash: bool = False,
                    check: bool = True) -> "PackedHashTable":
        target = torch.device(device) if device is not None else coords.device
        coords = coords.contiguous().to(dtype=torch.int32, device=target)
        n = coords.shape[0]
        cap = capacity if capacity is not None else max(16, n * 2)
        obj = cls(capacity=cap, device=target, use_double_hash=use_double_hash)
        obj.insert(coords, check=check)
        return obj

    def search(self, query_coords: Tensor, mode: SearchMode = SearchMode.LINEAR) -> Tensor:
        """int32 (M,) original insertion index, -1 when absent."""
        assert self._keys is not None, "Call insert() first"
        assert query_coords.ndim == 2 and query_coords.shape[1] == 4
        q = query_coords.contiguous().to(dtype=torch.int32, device=self.device)
        return _ops.hash_search(self._keys, self._values, q)

    @property
    def unique_index(self) -> Tensor:
        """Sorted indices of the first occurrence of every distinct coordinate."""
        indices = self.search(self._coords)
        return torch.unique(indices[indices != -1])
