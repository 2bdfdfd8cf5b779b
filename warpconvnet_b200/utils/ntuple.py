# SPDX-License-Identifier: Apache-2.0
"""``ntuple`` helper (same contract as warpconvnet/utils/ntuple.py:10-21)."""
from itertools import repeat
from typing import List, Tuple, Union

import torch


def ntuple(x: Union[int, List[int], Tuple[int, ...], torch.Tensor], ndim: int) -> Tuple[int, ...]:
    if isinstance(x, int):
        x = tuple(repeat(x, ndim))
    elif isinstance(x, list):
        x = tuple(x)
    elif isinstance(x, torch.Tensor):
        x = tuple(int(v) for v in x.view(-1).cpu().tolist())
    assert isinstance(x, tuple) and len(x) == ndim, x
    return x
