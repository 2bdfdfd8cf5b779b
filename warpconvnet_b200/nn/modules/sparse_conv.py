# SPDX-License-Identifier: Apache-2.0
"""Sparse convolution modules on ``Voxels``.

API-compatible with the reference's ``SpatiallySparseConv`` / ``SparseConv2d`` / ``SparseConv3d``
(warpconvnet/nn/modules/sparse_conv.py:31-391): the same constructor keywords, ``weight`` of shape
[K, Cin, Cout] (dense) or [K, G, Cin/G, Cout/G] (grouped), ``bias`` of shape [Cout], and the same
kaiming-uniform bounds scaled by sqrt(num_spatial_dims) (:182-217), so state dicts interchange.
All algorithm selectors map to the one tcgen05 backend of this package.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from warpconvnet_b200.geometry.types.voxels import Voxels
from warpconvnet_b200.nn.functional import sparse_conv as _F
from warpconvnet_b200.nn.modules.base_module import BaseSpatialModule
from warpconvnet_b200.utils.ntuple import ntuple

STRIDED_CONV_MODE = _F.STRIDED_CONV_MODE
SPARSE_CONV_AB_ALGO_MODE = _F.SPARSE_CONV_AB_ALGO_MODE
SPARSE_CONV_ATB_ALGO_MODE = _F.SPARSE_CONV_ATB_ALGO_MODE

# keyword -> (enum used to parse strings); None selects the package's only backend
_ALGO_KEYS = {"fwd_algo": SPARSE_CONV_AB_ALGO_MODE, "dgrad_algo": SPARSE_CONV_AB_ALGO_MODE,
              "wgrad_algo": SPARSE_CONV_ATB_ALGO_MODE}


def _select_algo(value, enum_cls):
    """Strings are parsed, enum members and lists pass through (the reference accepts a list of
    candidate names, sparse_conv.py:130-137)."""
    if value is None:
        return enum_cls.TCGEN05
    return enum_cls(value) if isinstance(value, str) else value


class SpatiallySparseConv(BaseSpatialModule):
    """Y[out] = bias + sum_k X[in_k(out)] @ weight[k] with the output coordinates decided by
    stride / transposed / generative exactly like the reference's functional."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride=1, dilation=1,
                 bias: bool = True, transposed: bool = False, generative: bool = False,
                 groups: int = 1, kernel_matmul_batch_size: int = 2,
                 num_spatial_dims: Optional[int] = 3, fwd_algo=None, dgrad_algo=None,
                 wgrad_algo=None, stride_mode=STRIDED_CONV_MODE.STRIDE_ONLY, order=None,
                 compute_dtype: Optional[torch.dtype] = None,
                 use_fp16_accum: Optional[bool] = None,
                 implicit_matmul_fwd_block_size: Optional[int] = None,
                 implicit_matmul_bwd_block_size: Optional[int] = None):
        super().__init__()
        for label, count in (("in_channels", in_channels), ("out_channels", out_channels)):
            if count % groups:
                raise ValueError(f"{label} ({count}) must be divisible by groups ({groups})")
        nd = num_spatial_dims
        window = dict(kernel_size=kernel_size, stride=stride, dilation=dilation)
        for key in window:
            setattr(self, key, ntuple(window[key], ndim=nd))
        chosen = dict(fwd_algo=fwd_algo, dgrad_algo=dgrad_algo, wgrad_algo=wgrad_algo)
        for key, enum_cls in _ALGO_KEYS.items():
            setattr(self, key, _select_algo(chosen[key], enum_cls))
        plain = dict(num_spatial_dims=nd, in_channels=in_channels, out_channels=out_channels,
                     groups=groups, transposed=transposed, generative=generative,
                     kernel_matmul_batch_size=kernel_matmul_batch_size, stride_mode=stride_mode,
                     order=order, compute_dtype=compute_dtype, use_fp16_accum=use_fp16_accum,
                     implicit_matmul_fwd_block_size=implicit_matmul_fwd_block_size,
                     implicit_matmul_bwd_block_size=implicit_matmul_bwd_block_size)
        for key, value in plain.items():
            setattr(self, key, value)
        # Extension (not a reference argument): when a BatchNorm follows this conv, set
        # ``conv.emit_bn_stats = True`` and the GEMM epilogue accumulates the per-channel sum / sum of
        # squares of the output it stores; warpconvnet_b200.nn.modules.BatchNorm picks them up
        # and skips its statistics pass over Y.
        self.emit_bn_stats = False

        volume = int(np.prod(self.kernel_size))
        shape = ((volume, in_channels, out_channels) if groups == 1
                 else (volume, groups, in_channels // groups, out_channels // groups))
        self.weight = nn.Parameter(torch.empty(shape))
        self.bias: Optional[nn.Parameter] = (nn.Parameter(torch.empty(out_channels))
                                             if bias else None)
        self.reset_parameters()

    # -- initialisation ---------------------------------------------------------------------
    def _calculate_fan_in_and_fan_out(self):
        volume = int(np.prod(self.kernel_size))
        return volume * (self.in_channels // self.groups), volume * (self.out_channels // self.groups)

    def _custom_kaiming_uniform_(self, tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
        fans = dict(zip(("fan_in", "fan_out"), self._calculate_fan_in_and_fan_out()))
        limit = (math.sqrt(self.num_spatial_dims) * nn.init.calculate_gain(nonlinearity, a)
                 / math.sqrt(fans[mode]))
        with torch.no_grad():
            return tensor.uniform_(-limit, limit)

    @torch.no_grad()
    def reset_parameters(self):
        self._custom_kaiming_uniform_(self.weight, a=math.sqrt(5),
                                      mode="fan_out" if self.transposed else "fan_in")
        if self.bias is not None:
            limit = 1.0 / math.sqrt(self._calculate_fan_in_and_fan_out()[0])
            self.bias.uniform_(-limit, limit)

    def __repr__(self):
        shown = [("in_channels", self.in_channels, None), ("out_channels", self.out_channels, None),
                 ("kernel_size", self.kernel_size, None),
                 ("stride", self.stride, (1,) * len(self.stride)),
                 ("dilation", self.dilation, (1,) * len(self.dilation)),
                 ("groups", self.groups, 1), ("transposed", self.transposed, False),
                 ("generative", self.generative, False)]
        body = ", ".join(f"{k}={v}" for k, v, default in shown if default is None or v != default)
        return f"{type(self).__name__}({body})"

    # -- forward ----------------------------------------------------------------------------
    def forward(self, input_sparse_tensor: Voxels,
                output_spatially_sparse_tensor: Optional[Voxels] = None):
        options = {key: getattr(self, key) for key in (
            "stride", "bias", "groups", "kernel_matmul_batch_size", "transposed", "generative",
            "fwd_algo", "dgrad_algo", "wgrad_algo", "stride_mode", "order", "compute_dtype",
            "use_fp16_accum", "implicit_matmul_fwd_block_size", "implicit_matmul_bwd_block_size")}
        return _F.spatially_sparse_conv(
            input_sparse_tensor=input_sparse_tensor, weight=self.weight,
            kernel_size=self.kernel_size, kernel_dilation=self.dilation,
            output_spatially_sparse_tensor=output_spatially_sparse_tensor,
            bn_stats=self.emit_bn_stats and self.training, **options)


def _with_dims(nd: int, name: str):
    """Subclass with ``num_spatial_dims`` fixed (SparseConv2d / SparseConv3d of the reference)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, bias=True,
                 transposed=False, generative: bool = False, groups: int = 1,
                 stride_mode=STRIDED_CONV_MODE.STRIDE_ONLY, fwd_algo=None, dgrad_algo=None,
                 wgrad_algo=None, kernel_matmul_batch_size: int = 2, order=None,
                 compute_dtype: Optional[torch.dtype] = None,
                 use_fp16_accum: Optional[bool] = None):
        SpatiallySparseConv.__init__(
            self, in_channels, out_channels, kernel_size, stride=stride, dilation=dilation,
            bias=bias, transposed=transposed, generative=generative, groups=groups,
            kernel_matmul_batch_size=kernel_matmul_batch_size, num_spatial_dims=nd,
            fwd_algo=fwd_algo, dgrad_algo=dgrad_algo, wgrad_algo=wgrad_algo,
            stride_mode=stride_mode, order=order, compute_dtype=compute_dtype,
            use_fp16_accum=use_fp16_accum)

    return type(name, (SpatiallySparseConv,),
                {"__init__": __init__, "__module__": __name__,
                 "__doc__": f"{nd}-D sparse convolution ({name} of the reference)."})


SparseConv2d = _with_dims(2, "SparseConv2d")
SparseConv3d = _with_dims(3, "SparseConv3d")
