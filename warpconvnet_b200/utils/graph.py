# SPDX-License-Identifier: Apache-2.0
"""Whole-step CUDA-graph capture for a path whose tensor SIZES depend on the data.

Everything on the sparse-conv path enqueues without a host synchronisation except the few
operations that create a new coordinate set (``stride_coords`` / ``expand_coords`` / ``unique``):
the number of output rows is only known on the device, and the host needs it to size every tensor
that follows. In eager mode those operations read the count back (one small D2H on a side stream).
Under stream capture nothing may be read back, so the sizes are taken from a ``SizeTape`` recorded
during an eager warm-up pass over the SAME coordinates:

    tape = SizeTape()
    with tape.record():
        step()                         # eager, sizes are read back and appended to the tape
    graph = torch.cuda.CUDAGraph()
    with tape.replay(), torch.cuda.graph(graph):
        step()                         # captured, sizes come from the tape in call order
    graph.replay()

The replayed graph is valid for as long as the coordinate sets keep their sizes, i.e. for a fixed
geometry (benchmark protocol, multi-epoch training on cached scenes, static-scene inference);
every entry is checked against the call signature (input rows, stride, offsets, batch items) and
the device-side totals are verified by ``SizeTape.verify()`` after a replay. With changing
geometry the step runs eagerly — still without blocking the compute stream.
(The reference cannot be captured at all: its kernel-map build synchronises >= 6 times per map,
SURVEY.md §3.1.)
"""
from __future__ import annotations

import contextlib
from typing import Any, List, Optional, Tuple

import torch

_ACTIVE: List["SizeTape"] = []


def active_tape() -> Optional["SizeTape"]:
    return _ACTIVE[-1] if _ACTIVE else None


class SizeTape:
    def __init__(self):
        self.entries: List[Tuple[Any, Any]] = []   # (signature, payload)
        self.mode: Optional[str] = None
        self._cursor = 0
        self._checks: List[Tuple[torch.Tensor, int]] = []  # (device total word, expected)

    # ---- used by the size-producing ops ----------------------------------------------------
    @property
    def recording(self) -> bool:
        return self.mode == "record"

    @property
    def replaying(self) -> bool:
        return self.mode == "replay"

    def append(self, signature, payload) -> None:
        self.entries.append((signature, payload))

    def next(self, signature):
        if self._cursor >= len(self.entries):
            raise RuntimeError(
                f"SizeTape exhausted at call {self._cursor} ({signature}): the captured step "
                "creates more coordinate sets than the recorded warm-up pass")
        sig, payload = self.entries[self._cursor]
        if sig != signature:
            raise RuntimeError(
                f"SizeTape mismatch at call {self._cursor}: recorded {sig}, replaying {signature}")
        self._cursor += 1
        return payload

    def expect(self, total_word: torch.Tensor, expected: int) -> None:
        """Remember a device word that must equal ``expected`` after every replay."""
        self._checks.append((total_word, int(expected)))

    # ---- user side --------------------------------------------------------------------------
    @contextlib.contextmanager
    def record(self):
        self.entries, self.mode, self._cursor = [], "record", 0
        _ACTIVE.append(self)
        try:
            yield self
        finally:
            _ACTIVE.pop()
            self.mode = None

    @contextlib.contextmanager
    def replay(self):
        self.mode, self._cursor, self._checks = "replay", 0, []
        _ACTIVE.append(self)
        try:
            yield self
        finally:
            _ACTIVE.pop()
            self.mode = None

    def verify(self) -> None:
        """After ``graph.replay()`` + synchronize: raises when a coordinate set produced another
        number of rows than the graph was captured for (the geometry changed)."""
        for word, expected in self._checks:
            got = int(word.item())
            if got != expected:
                raise RuntimeError(
                    f"captured graph replayed on a different geometry: a coordinate set has {got} "
                    f"rows, the graph was captured for {expected}")


def capture_step(step_fn, warmup: int = 1, pool=None):
    """Records the sizes of ``step_fn`` eagerly (``warmup`` passes, the last one on tape), captures
    it into a CUDA graph and returns ``(graph, tape, outputs_of_the_captured_call)``."""
    tape = SizeTape()
    for _ in range(max(warmup - 1, 0)):
        step_fn()
    with tape.record():
        step_fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with tape.replay():
        with torch.cuda.graph(graph, pool=pool, capture_error_mode="relaxed"):
            out = step_fn()
    return graph, tape, out
