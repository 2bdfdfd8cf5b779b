# SPDX-License-Identifier: Apache-2.0
"""Scene-sharded data parallelism for the sparse-conv path.

The path shards by batch item: the batch index is part of the packed hash key, so no kernel-map
pair crosses scenes and forward / dgrad are row-local. Only the weight gradients sum over all
scenes, so the single collective is ONE all-reduce of the fp32 wgrad buffers per step (NCCL over
NVLink on the GPU box, gloo in the CPU tests). The reference has no collective code at all
(grep over warpconvnet/ is empty); users wrap DDP. This module is the minimal
equivalent: a flat bucket so every layer's dW lands in one contiguous buffer and the whole
model needs one NCCL launch.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_scenes(num_scenes: int, rank: int, world_size: int) -> List[int]:
    """Scenes owned by `rank`: {s : s mod world_size == rank} (round robin keeps voxel counts
    balanced when scene sizes are i.i.d.)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, num_scenes, world_size))


class FlatGradBucket:
    """One contiguous fp32 buffer holding the gradients of `params`; ``param.grad`` are views
    into it, so the wgrad kernels' output (accumulated by autograd into .grad) is reduced with a
    single collective and no packing copy.

    ``optimizer.zero_grad(set_to_none=True)`` (torch's default) drops the views: autograd then
    allocates fresh ``.grad`` tensors. ``all_reduce`` therefore re-binds first — a gradient that no
    longer aliases the buffer is copied into its slot and ``.grad`` is pointed back at the slot —
    so the collective never reduces a stale buffer. Use ``bucket.zero()`` (or
    ``zero_grad(set_to_none=False)``) between steps to keep the views and skip that copy."""

    def __init__(self, params: Iterable[torch.nn.Parameter], peer: bool = False,
                 peer_ctas: int = 64, sections: int = 1):
        """``peer``: keep the buffer in NVLink peer-mapped memory and reduce it with this
        library's own kernel (``PeerAllReduce``) instead of an NCCL collective. ``sections`` > 1
        (peer only): the buffer is cut into that many pieces and the reduction of a piece is
        launched — on a side stream, by a post-accumulate hook — as soon as every gradient that
        overlaps it has been produced, so it runs under the rest of backward (what DDP does with
        its buckets); ``all_reduce()`` launches what is left and joins."""
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.peer = None
        self._sections: List[tuple] = []
        if peer and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            n_sec = max(1, int(sections))
            self.peer = PeerAllReduce(total, dev, n_ctas=peer_ctas, n_sections=n_sec)
            self.flat = self.peer.buffer
            if n_sec > 1:
                self._setup_sections(n_sec, total)
        else:
            self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self._slots = []
        off = 0
        for p in self.params:
            n = p.numel()
            self._slots.append(self.flat[off:off + n].view_as(p))
            off += n
        self.bind()

    # ---- sections: pieces of the peer buffer reduced while backward is still running ------------
    def _setup_sections(self, n_sec: int, total: int) -> None:
        n_pad = self.peer.n
        bounds = [min(n_pad, (n_pad * k // n_sec) // 4 * 4) for k in range(n_sec)] + [n_pad]
        starts, off = [], 0
        for p in self.params:
            starts.append(off)
            off += p.numel()
        for k in range(n_sec):
            lo, hi = bounds[k], bounds[k + 1]
            members = [i for i, p in enumerate(self.params)
                       if starts[i] < hi and starts[i] + p.numel() > lo]
            self._sections.append((lo, hi - lo, members))
        self._of_param = [[k for k, (_, _, m) in enumerate(self._sections) if i in m]
                          for i in range(len(self.params))]
        self._ar_stream = torch.cuda.Stream(device=self.flat.device)
        self._average = True
        self._reset_sections()
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(lambda _p, i=i: self._grad_ready(i))

    def _reset_sections(self) -> None:
        self._seen = [False] * len(self.params)
        self._left = [len(m) for _, _, m in self._sections]
        self._launched = [False] * len(self._sections)

    def _launch_section(self, k: int) -> None:
        lo, length, _ = self._sections[k]
        self._launched[k] = True
        cur = torch.cuda.current_stream(self.flat.device)
        self._ar_stream.wait_stream(cur)
        with torch.cuda.stream(self._ar_stream):
            self.peer.all_reduce_section_(k, lo, length, average=self._average)

    def _grad_ready(self, i: int) -> None:
        p = self.params[i]
        if self._seen[i] or p.grad is None or p.grad.data_ptr() != self._slots[i].data_ptr():
            return  # accumulated twice (shared parameter) or not bound: all_reduce() picks it up
        self._seen[i] = True
        for k in self._of_param[i]:
            self._left[k] -= 1
            if self._left[k] == 0 and not self._launched[k]:
                self._launch_section(k)

    def bind(self) -> int:
        """Point every ``param.grad`` at its slot; returns how many had to be re-bound."""
        rebound = 0
        for p, slot in zip(self.params, self._slots):
            g = p.grad
            if g is not None and g.data_ptr() == slot.data_ptr() and g.dtype == torch.float32:
                continue
            if g is None:
                slot.zero_()
            else:
                slot.copy_(g)
            p.grad = slot
            rebound += 1
        return rebound

    def set_average(self, average: bool) -> None:
        """What the section launches made during backward apply (default: the mean)."""
        self._average = bool(average)

    def zero(self) -> None:
        self.flat.zero_()
        self.bind()

    def all_reduce(self, average: bool = False, async_op: bool = False):
        """Sum (or mean) over all ranks. With ``average`` the buffer is pre-scaled by 1 / world
        BEFORE the collective, so the result is already the mean when an ``async_op`` handle
        completes."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            self.bind()
            return None
        rebound = self.bind()
        if self.peer is not None and self._sections:
            # sections launched by the hooks reduced what their gradients held at that time: if a
            # gradient had to be re-bound (copied into its slot just now) the early launches were
            # made with the wrong ``average`` or stale data — not supported, say so
            if rebound and any(self._launched):
                raise RuntimeError("FlatGradBucket(sections > 1) needs .grad to stay bound during "
                                   "backward: use bucket.zero() instead of zero_grad(set_to_none=True)")
            if average != self._average and any(self._launched):
                raise RuntimeError("FlatGradBucket(sections > 1): all_reduce(average=...) must match "
                                   "bucket.set_average(...) used by the early launches")
            self._average = average
            for k in range(len(self._sections)):
                if not self._launched[k]:
                    self._launch_section(k)
            torch.cuda.current_stream(self.flat.device).wait_stream(self._ar_stream)
            self._reset_sections()
            return None
        if self.peer is not None:  # stream-ordered kernel: nothing to wait for
            self.peer.all_reduce_(average=average)
            return None
        if average:
            self.flat.mul_(1.0 / dist.get_world_size())
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)


class PeerAllReduce:
    """In-place sum all-reduce of one fp32 buffer over NVLink peer memory with this library's own
    kernel (``wcn_peer_allreduce_f32``, csrc/peer_allreduce.cu) instead of an NCCL collective:
    two flag barriers and a two-shot reduce, a few microseconds for the 1.77 MB dW of a
    128 -> 128 layer, launched on the caller's stream (so it is captured into a CUDA graph with
    the rest of the step) and small enough (32 CTAs of 128 threads, no shared memory) to run beside the dgrad kernel.

    ``buffer`` is the tensor to produce the gradients in (``sparse_conv_wgrad(..., out=buffer)``
    or the ``flat`` storage of a ``FlatGradBucket``): it is allocated as symmetric memory —
    ``torch.distributed._symmetric_memory`` is used for the allocation and the handle exchange
    only. Single node; every rank constructs it and calls ``all_reduce_()`` in the same order."""

    def __init__(self, numel: int, device, group=None, n_ctas: int = 32, n_sections: int = 1):
        """``n_sections`` > 1: the buffer can also be reduced piecewise (``all_reduce_section_``),
        every section through its own flag block, so several sections may be in flight at once
        (gradient buckets launched while backward is still running)."""
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        from ._lib import lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerAllReduce needs an initialised process group")
        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n = (int(numel) + 3) // 4 * 4
        self.n_ctas = int(n_ctas)
        words = int(lib.wcn_peer_allreduce_flag_words())
        self._words = words
        self.n_sections = max(1, int(n_sections))
        self._timeout_word = int(lib.wcn_peer_allreduce_timeout_word())
        # the allocation is local and may fail on one rank only; the rendezvous below is collective:
        # agree first, so that every rank either goes on or raises (callers fall back to NCCL)
        err = None
        try:
            self._data = symm_mem.empty(self.n, dtype=torch.float32, device=device)
            self._flags = symm_mem.empty(words * self.n_sections, dtype=torch.int32, device=device)
            self._data.zero_()
            self._flags.zero_()
        except Exception as exc:  # pragma: no cover - depends on the driver / allocator state
            err = exc
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            raise RuntimeError("symmetric memory allocation failed on some rank"
                               + (f": {err}" if err is not None else ""))
        try:
            self._h_data = symm_mem.rendezvous(self._data, group)
            self._h_flags = symm_mem.rendezvous(self._flags, group)
        except TypeError:  # older signature: group name
            self._h_data = symm_mem.rendezvous(self._data, group.group_name)
            self._h_flags = symm_mem.rendezvous(self._flags, group.group_name)
        arr = ctypes.c_void_p * self.world
        self._bufs = arr(*[int(p) for p in self._h_data.buffer_ptrs])
        self._flag_ptrs = arr(*[int(p) for p in self._h_flags.buffer_ptrs])
        self.buffer = self._data[:int(numel)]
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every rank's flags are zero before the first kernel touches them

    def timeouts(self) -> int:
        """Barrier waits that gave up (10 s) since construction — synchronises; non-zero means a
        peer never arrived and the buffer is not a valid sum."""
        return int(self._flags.view(self.n_sections, self._words)[:, self._timeout_word].sum().item())

    def all_reduce_(self, average: bool = False) -> torch.Tensor:
        """The whole buffer (through the flag block of section 0)."""
        self.all_reduce_section_(0, 0, self.n, average)
        return self.buffer

    def all_reduce_section_(self, section: int, start: int, length: int, average: bool = False) -> None:
        """Elements [start, start + length) (both multiples of 4) through flag block ``section``.
        Every rank issues the same sections; a section must not be issued again before its
        previous reduction has been ordered before the new one on this rank (same stream, or a
        stream dependency)."""
        import ctypes

        from ._lib import check, lib
        if not (0 <= section < self.n_sections) or start % 4 or length % 4 or start < 0 \
                or start + length > self.n:
            raise ValueError("bad section / range")
        if length == 0:
            return
        arr = ctypes.c_void_p * self.world
        bufs = arr(*[int(p) + 4 * start for p in self._bufs])
        flags = arr(*[int(p) + 4 * self._words * section for p in self._flag_ptrs])
        scale = ctypes.c_float(1.0 / self.world if average else 1.0)   # applied inside the kernel
        check(lib.wcn_peer_allreduce_f32(bufs, flags, self.rank, self.world, length, scale,
                                         self.n_ctas, torch.cuda.current_stream().cuda_stream),
              "peer_allreduce")


def all_reduce_wgrad(tensors: Sequence[torch.Tensor], average: bool = False) -> None:
    """Sum a list of fp32 weight-gradient tensors over all ranks with one flattened collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat.div_(dist.get_world_size())
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
