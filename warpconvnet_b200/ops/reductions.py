# SPDX-License-Identifier: Apache-2.0
"""CSR row reductions for ``PointConv`` (warpconvnet/ops/reductions.py:36-75; the reference uses
``torch_scatter.segment_csr``, which is not a dependency here).

kNN neighbourhoods have a fixed row length, so the common case is a plain ``[M, k, C]`` view
reduced along ``k`` — exact, fully differentiable, no scatter. Ragged rows (radius search) go
through ``torch.segment_reduce``. Max / min backward on exact ties: the fixed-k path routes the
gradient to the first maximal element like ``segment_csr``; the ragged path follows
``torch.segment_reduce`` (the reference warns about that difference, reductions.py:56-61)."""
from enum import Enum
from typing import Literal

import torch
from torch import Tensor


class REDUCTIONS(Enum):
    MIN = "min"
    MAX = "max"
    MEAN = "mean"
    SUM = "sum"
    MUL = "mul"
    VAR = "var"
    STD = "std"
    RANDOM = "random"


REDUCTION_TYPES_STR = Literal["min", "max", "mean", "sum", "mul", "var", "std", "random"]


def _uniform_k(row_offsets: Tensor, n: int):
    m = row_offsets.numel() - 1
    if m <= 0 or n % m != 0:
        return None
    k = n // m
    # row_splits built by RealSearchResult for [M, K] tensors are exactly arange(0, M*K+1, K)
    return k if getattr(row_offsets, "_wcn_uniform_k", None) == k else None


def _segment(features: Tensor, row_offsets: Tensor, reduce: str) -> Tensor:
    lengths = row_offsets.diff().to(features.device)
    return torch.segment_reduce(features, reduce, lengths=lengths, axis=0, unsafe=True)


def row_reduction(features: Tensor, row_offsets: Tensor, reduction, eps: float = 1e-6) -> Tensor:
    if isinstance(reduction, str):
        reduction = REDUCTIONS(reduction)
    n = features.shape[0]
    k = _uniform_k(row_offsets, n)
    if k is not None:
        f = features.view(-1, k, features.shape[-1])
        if reduction == REDUCTIONS.MEAN:
            return f.mean(dim=1)
        if reduction == REDUCTIONS.SUM:
            return f.sum(dim=1)
        if reduction == REDUCTIONS.MAX:
            return f.max(dim=1).values
        if reduction == REDUCTIONS.MIN:
            return f.min(dim=1).values
        if reduction == REDUCTIONS.MUL:
            return f.prod(dim=1)
        if reduction == REDUCTIONS.VAR:
            return (f * f).mean(dim=1) - f.mean(dim=1) ** 2
        if reduction == REDUCTIONS.STD:
            return torch.sqrt((f * f).mean(dim=1) - f.mean(dim=1) ** 2 + eps)
        if reduction == REDUCTIONS.RANDOM:
            pick = torch.randint(0, k, (f.shape[0],), device=f.device)
            return f[torch.arange(f.shape[0], device=f.device), pick]
        raise ValueError(f"Invalid reduction: {reduction}")
    assert n == int(row_offsets[-1]), \
        f"Features length {n} must match the last row split {int(row_offsets[-1])}"
    if reduction in (REDUCTIONS.MIN, REDUCTIONS.MAX, REDUCTIONS.MEAN, REDUCTIONS.SUM):
        return _segment(features, row_offsets, reduction.value)
    if reduction == REDUCTIONS.MUL:
        return _segment(features, row_offsets, "prod")
    if reduction in (REDUCTIONS.VAR, REDUCTIONS.STD):
        mean = _segment(features, row_offsets, "mean")
        var = _segment(features * features, row_offsets, "mean") - mean ** 2
        return var if reduction == REDUCTIONS.VAR else torch.sqrt(var + eps)
    if reduction == REDUCTIONS.RANDOM:
        num = row_offsets.diff()
        pick = (torch.rand(len(num), device=num.device) * num).floor().long() + row_offsets[:-1]
        return features[pick.to(features.device)]
    raise ValueError(f"Invalid reduction: {reduction}")
