# SPDX-License-Identifier: Apache-2.0
"""Kernel-map and neighbour-search containers.

``IntSearchResult`` keeps the reference's fields and methods
(warpconvnet/geometry/coords/search/search_results.py:55-202): ``in_maps``, ``out_maps`` on the
device, ``offsets`` on the CPU, ``identity_map_index``, ``__getitem__`` / ``__len__`` / ``numel`` /
``clone`` / ``get_batch``. On top of that it lazily caches what the tensor-core kernels need:
the dense pair table, its reverse, the device copy of ``offsets`` and the mask-sorted tile plans.
"""
from __future__ import annotations

from typing import List, Literal, Optional, Tuple

import torch
from torch import Tensor

from warpconvnet_b200 import _ops


# Deferred kernel maps whose (offsets, status) copy is still in flight. The conv path never reads
# them on the host, so without this list an out-of-range coordinate or a full hash table (status
# word) would go unnoticed and training would continue on a wrong kernel map. Every map build and
# every conv backward polls the list WITHOUT blocking; entries whose copy has landed are checked
# (ValueError / RuntimeError like the reference raises at build time) and dropped, and the oldest
# entries are waited for once more than _MAX_PENDING are outstanding, so a failure surfaces within
# the step that caused it.
_PENDING_STATUS: List[tuple] = []
_MAX_PENDING = 256


class _PinnedArena:
    """ONE pinned allocation, cut into slots that the kernel-map builder takes round robin for its
    asynchronous (offsets, status) read-back. ``torch.empty(pin_memory=True)`` per map looked free
    (the caching host allocator reuses blocks) but is not: a block is only reusable once the copy
    that wrote it has completed, so whenever the host runs further ahead of the GPU than ever
    before, the allocator falls through to ``cudaHostAlloc`` — 10-30 ms of blocked host thread in
    the middle of a step (the sporadic stalls of back-to-back eager steps: single steps of 4-33 ms
    in a 1 ms/step loop, profiles/r2x_e2e_stalls.md)."""
    SLOTS, WIDTH = 1024, 128

    def __init__(self):
        self.buf = torch.empty((self.SLOTS, self.WIDTH), dtype=torch.int32, pin_memory=True)
        self.events = [None] * self.SLOTS
        self.gen = [0] * self.SLOTS
        self.next = 0

    def acquire(self, n: int):
        i = self.next
        self.next = (i + 1) % self.SLOTS
        ev = self.events[i]
        if ev is not None and not ev.query():   # the GPU is 1024 maps behind: wait for it
            ev.synchronize()
        self.gen[i] += 1
        return i, self.gen[i], self.buf[i, :n]


_ARENA: Optional[_PinnedArena] = None


class HostCopy:
    """(offsets[K+1], status) of one deferred kernel map on its way to the host: a slot of the
    pinned arena plus the event recorded behind the copy, or — built under CUDA-graph capture,
    or when the slot has been handed out again since — the device tensor itself (read
    synchronously when somebody asks)."""

    def __init__(self, meta_dev: Tensor, asynchronous: bool):
        global _ARENA
        self.meta_dev = meta_dev
        self.value: Optional[Tensor] = None
        self.event = None
        self.slot = -1
        self.gen = 0
        n = meta_dev.numel()
        if asynchronous and n <= _PinnedArena.WIDTH:
            if _ARENA is None:
                _ARENA = _PinnedArena()
            self.slot, self.gen, view = _ARENA.acquire(n)
            view.copy_(meta_dev, non_blocking=True)
            self.view = view
            self.event = torch.cuda.Event()
            self.event.record(torch.cuda.current_stream(meta_dev.device))
            _ARENA.events[self.slot] = self.event

    def ready(self) -> bool:
        return self.value is not None or (self.event is not None and self.event.query())

    def read(self) -> Tensor:
        """CPU int32 [K + 2]; waits for the copy (or reads the device tensor) on first use."""
        if self.value is None:
            if self.event is not None:
                self.event.synchronize()
                if _ARENA.gen[self.slot] == self.gen:
                    self.value = self.view.clone()
            if self.value is None:              # no slot / slot reused since: synchronous read
                self.value = self.meta_dev.cpu()
            self.meta_dev = None
        return self.value


def _register_pending(host: "HostCopy", on_ready) -> None:
    if host.event is not None and on_ready is not None:
        _PENDING_STATUS.append((host, on_ready))


def check_pending_kernel_maps(block: bool = False) -> int:
    """Raises the deferred hash-table errors of kernel maps built so far whose status has reached
    the host (``block=True``: waits for all of them — call it at the end of a step when a
    guaranteed check is wanted). Returns the number still in flight."""
    while _PENDING_STATUS:
        host, on_ready = _PENDING_STATUS[0]
        if not (block or len(_PENDING_STATUS) > _MAX_PENDING or host.ready()):
            break
        _PENDING_STATUS.pop(0)
        on_ready(int(host.read()[-1]))
    return len(_PENDING_STATUS)


class RealSearchResult:
    """CSR neighbour lists (search_results.py:13-52)."""

    def __init__(self, *args):
        if len(args) == 2:
            self.neighbor_indices = args[0].long()
            self.neighbor_row_splits = args[1].long()
        elif len(args) == 1:
            assert isinstance(args[0], Tensor) and args[0].ndim == 2
            M, K = args[0].shape
            self.neighbor_indices = args[0].long().reshape(-1)
            self.neighbor_row_splits = torch.arange(0, M * K + 1, K, device=args[0].device,
                                                    dtype=torch.long)
            self.neighbor_row_splits._wcn_uniform_k = K  # fixed row length (kNN): see reductions
        else:
            raise ValueError("RealSearchResult must be initialized with 1 or 2 arguments")
        self.neighbor_distances = None

    def to(self, device):
        self.neighbor_indices = self.neighbor_indices.to(device)
        self.neighbor_row_splits = self.neighbor_row_splits.to(device)
        return self

    def __repr__(self):
        return (f"{self.__class__.__name__}(neighbor_indices={tuple(self.neighbor_indices.shape)}, "
                f"neighbor_row_splits={tuple(self.neighbor_row_splits.shape)})")


class IntSearchResult:
    """Kernel map in CSR form. Two construction modes:

    * eager (the reference's constructor): ``in_maps``, ``out_maps`` of exact length L and
      ``offsets`` (copied to the CPU here, like search_results.py:73-76);
    * deferred (``_from_device``, used by ``generate_kernel_map``): the maps live in upper-bound
      sized device buffers, ``offsets`` stay on the device and an asynchronous copy to pinned host
      memory is in flight. Nothing synchronises until ``offsets`` / ``in_maps`` / ``out_maps`` /
      ``len(in_maps)`` are actually read on the host; the conv kernels only use the raw device
      buffers, so a whole fwd+bwd step enqueues without a single host sync (the reference needs
      >= 6 per map).
    """

    def __init__(self, in_maps: Tensor, out_maps: Tensor, offsets: Tensor,
                 identity_map_index: Optional[int] = None):
        offsets_cpu = offsets.cpu()
        assert len(in_maps) == len(out_maps) == int(offsets_cpu[-1])
        self._in_maps: Optional[Tensor] = in_maps
        self._out_maps: Optional[Tensor] = out_maps
        self._offsets: Optional[Tensor] = offsets_cpu
        self._in_buf, self._out_buf = in_maps, out_maps
        self._pending = None
        self._init_common(identity_map_index, offsets if offsets.is_cuda else None)

    def _init_common(self, identity_map_index, offsets_dev):
        self.identity_map_index = identity_map_index
        # lazily built device-side state for the sm_100a kernels
        self._offsets_dev: Optional[Tensor] = offsets_dev
        self._pair_table: Optional[Tensor] = None      # [K, n_out] -> input row
        self._rev_pair_table: Optional[Tensor] = None  # [K, n_in]  -> output row
        self._mask_keys: Optional[Tensor] = None
        self._fwd_plan: Optional[_ops.TilePlan] = None
        self._bwd_plan: Optional[_ops.TilePlan] = None
        self._n_in: Optional[int] = None
        self._n_out: Optional[int] = None
        # True when in/out coordinates are the same set and the kernel is odd: then
        # rev_pair_table[k] == pair_table[K-1-k] and dgrad reuses the forward tile plan.
        self._symmetric = False

    @classmethod
    def _from_device(cls, in_buf: Tensor, out_buf: Tensor, offsets_dev: Tensor, host: "HostCopy",
                     on_ready, identity_map_index: Optional[int]) -> "IntSearchResult":
        """host: the (offsets, status) read-back in flight; on_ready(status): raises the deferred
        hash-table errors."""
        self = cls.__new__(cls)
        self._in_maps = self._out_maps = self._offsets = None
        self._in_buf, self._out_buf = in_buf, out_buf
        self._pending = (host, on_ready)
        _register_pending(host, on_ready)
        self._init_common(identity_map_index, offsets_dev)
        return self

    def _resolve(self) -> None:
        if self._pending is None:
            return
        host, on_ready = self._pending
        self._pending = None
        for i, entry in enumerate(_PENDING_STATUS):  # this map is checked right here
            if entry[0] is host:
                del _PENDING_STATUS[i]
                break
        meta = host.read()
        if on_ready is not None:
            on_ready(int(meta[-1]))
        self._offsets = meta[:-1].clone()
        n = int(self._offsets[-1])
        self._in_maps = self._in_buf[:n]
        self._out_maps = self._out_buf[:n]

    @property
    def in_maps(self) -> Tensor:
        self._resolve()
        return self._in_maps

    @property
    def out_maps(self) -> Tensor:
        self._resolve()
        return self._out_maps

    @property
    def offsets(self) -> Tensor:
        """CPU offsets[K+1] (synchronises on first access in deferred mode)."""
        self._resolve()
        return self._offsets

    def validate(self) -> "IntSearchResult":
        """Force the deferred checks (out-of-range coordinates, table full) now."""
        self._resolve()
        return self

    # ---- reference API ---------------------------------------------------------------------
    @torch.no_grad()
    def __getitem__(self, idx: int) -> Tuple[Tensor, Tensor]:
        start, end = int(self.offsets[idx]), int(self.offsets[idx + 1])
        return self.in_maps[start:end], self.out_maps[start:end]

    @torch.no_grad()
    def get_batch(self, start_idx: int, end_idx: int,
                  out_format: Literal["list", "tensor"] = "list"):
        in_maps = [self[i][0] for i in range(start_idx, end_idx)]
        out_maps = [self[i][1] for i in range(start_idx, end_idx)]
        if out_format == "list":
            return in_maps, out_maps
        if out_format == "tensor":
            width = max(len(m) for m in in_maps)
            it = torch.full((len(in_maps), width), -1, device=self.in_maps.device, dtype=torch.int64)
            ot = torch.full((len(in_maps), width), -1, device=self.in_maps.device, dtype=torch.int64)
            for i, (a, b) in enumerate(zip(in_maps, out_maps)):
                it[i, : len(a)] = a
                ot[i, : len(b)] = b
            return it, ot
        raise ValueError(f"Invalid output format: {out_format}")

    def __len__(self):
        return len(self.offsets) - 1

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __repr__(self):
        return f"{self.__class__.__name__}(len={len(self)}, iden_map={self.identity_map_index})"

    def numel(self, i: int) -> int:
        return int(self.offsets[i + 1] - self.offsets[i])

    def clone(self):
        return IntSearchResult(self.in_maps.clone(), self.out_maps.clone(), self.offsets.clone(),
                               self.identity_map_index)

    @torch.no_grad()
    def to_csr(self) -> Tuple[Tensor, Tensor, Tensor]:
        """(in rows grouped by output row, the output rows that have pairs, CPU row offsets) —
        search_results.py:147-174. A stable sort keeps the offset order inside every row, so the
        result is deterministic (the reference's ``torch.sort`` is not stable)."""
        out_sorted, perm = torch.sort(self.out_maps.long(), stable=True)
        rows, counts = torch.unique_consecutive(out_sorted, return_counts=True)
        offsets = torch.zeros(len(rows) + 1, dtype=torch.int64)
        offsets[1:] = torch.cumsum(counts.cpu(), dim=0)
        return self.in_maps[perm], rows, offsets

    @property
    def device(self):
        return self._in_buf.device

    @torch.no_grad()
    def neighbor_count_per_output(self, num_out: int) -> Tensor:
        counts = torch.zeros(num_out, dtype=torch.long, device=self.out_maps.device)
        counts.scatter_add_(0, self.out_maps.long(), torch.ones_like(self.out_maps, dtype=torch.long))
        return counts

    # ---- device-side plans -----------------------------------------------------------------
    @property
    def offsets_dev(self) -> Tensor:
        if self._offsets_dev is None or self._offsets_dev.device != self._in_buf.device:
            self._offsets_dev = self.offsets.to(device=self._in_buf.device, dtype=torch.int32)
        if self._offsets_dev.dtype != torch.int32:
            self._offsets_dev = self._offsets_dev.int()
        return self._offsets_dev

    @torch.no_grad()
    def pair_table(self, n_out: int) -> Tensor:
        if self._pair_table is None:
            self._pair_table = _ops.csr_to_pair_table(self._in_buf, self._out_buf,
                                                      self.offsets_dev, n_out)
        return self._pair_table

    @torch.no_grad()
    def rev_pair_table(self, n_in: int) -> Tensor:
        if self._rev_pair_table is None:
            self._rev_pair_table = _ops.csr_to_pair_table(self._out_buf, self._in_buf,
                                                          self.offsets_dev, n_in)
        return self._rev_pair_table

    @torch.no_grad()
    def fwd_plan(self, n_out: int) -> _ops.TilePlan:
        """Tile plan of the forward (output-stationary) pass over the n_out output rows."""
        if self._fwd_plan is None:
            self._fwd_plan = _ops.build_tile_plan(self.pair_table(n_out), self._mask_keys)
            self._mask_keys = None
        return self._fwd_plan

    @torch.no_grad()
    def bwd_plan(self, n_in: int):
        """(plan, kflip) of the dgrad pass over the n_in input rows."""
        if self._symmetric:
            return self.fwd_plan(n_in), True
        if self._bwd_plan is None:
            self._bwd_plan = _ops.build_tile_plan(self.rev_pair_table(n_in))
        return self._bwd_plan, False

    def transposed_view(self) -> "IntSearchResult":
        """in/out swapped (helper.py:486-497), sharing every cached table with roles swapped."""
        t = IntSearchResult.__new__(IntSearchResult)
        t._in_maps, t._out_maps, t._offsets = self._out_maps, self._in_maps, self._offsets
        t._in_buf, t._out_buf = self._out_buf, self._in_buf
        t._pending = self._pending
        t._init_common(None, self._offsets_dev)
        t._pair_table, t._rev_pair_table = self._rev_pair_table, self._pair_table
        t._fwd_plan, t._bwd_plan = self._bwd_plan, self._fwd_plan
        t._n_in, t._n_out = self._n_out, self._n_in
        return t
