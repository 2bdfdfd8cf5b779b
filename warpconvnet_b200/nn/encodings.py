# SPDX-License-Identifier: Apache-2.0
"""Sinusoidal positional encoding (warpconvnet/nn/encodings.py:33-62,
nn/functional/encodings.py:12-80)."""
import math

import torch
import torch.nn as nn
from torch import Tensor


def get_freqs(num_freqs: int, data_range: float = 2.0, device=None) -> Tensor:
    freqs = 2 ** torch.arange(start=0, end=num_freqs, device=device)
    return (2 * math.pi / data_range) * freqs


def sinusoidal_encoding(x: Tensor, num_channels=None, data_range=None, encoding_axis: int = -1,
                        freqs: Tensor = None, concat_input: bool = False) -> Tensor:
    """``[..., D] -> [..., D * num_channels (+ D)]`` (nn/functional/encodings.py:28-80, same
    keywords: either ``freqs`` or ``num_channels`` + ``data_range``)."""
    assert encoding_axis == -1, "Only encoding_axis=-1 is supported at the moment"
    if freqs is None:
        assert num_channels is not None and data_range is not None, \
            "num_channels and data_range must be provided if freqs are not given"
        assert num_channels % 2 == 0, f"num_channels must be even for sin/cos, got {num_channels}"
        freqs = get_freqs(num_channels // 2, data_range, device=x.device)
    x = x.unsqueeze(-1)
    fx = x * freqs.reshape((1,) * (x.dim() - 1) + freqs.shape)
    parts = [fx.cos(), fx.sin()] + ([x] if concat_input else [])
    return torch.cat(parts, dim=-1).flatten(start_dim=-2)


class SinusoidalEncoding(nn.Module):
    def __init__(self, num_channels: int, data_range: float = 2.0, concat_input: bool = True):
        super().__init__()
        assert num_channels % 2 == 0, f"num_channels must be even for sin/cos, got {num_channels}"
        self.num_channels = num_channels
        self.concat_input = concat_input
        self.register_buffer("freqs", get_freqs(num_channels // 2, data_range))

    def forward(self, x: Tensor) -> Tensor:
        return sinusoidal_encoding(x, freqs=self.freqs, concat_input=self.concat_input)
