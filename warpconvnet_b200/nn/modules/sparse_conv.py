# SPDX-License-Identifier: Apache-2.0
"""``SparseConv3d`` / ``SparseConv2d`` / ``SpatiallySparseConv`` modules
(drop-in for warpconvnet/nn/modules/sparse_conv.py:31-391: same constructor arguments, parameter
names, shapes ``weight[K,Cin,Cout]`` / ``[K,G,Cin/G,Cout/G]``, ``bias[Cout]`` and the same
kaiming-uniform initialisation with the sqrt(num_spatial_dims) bound, :182-217)."""
from __future__ import annotations

import math
from typing import Literal, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn
from torch.nn import init
from torch.nn.init import calculate_gain

from warpconvnet_b200.geometry.types.voxels import Voxels
from warpconvnet_b200.nn.functional.sparse_conv import (SPARSE_CONV_AB_ALGO_MODE,
                                                        SPARSE_CONV_ATB_ALGO_MODE,
                                                        STRIDED_CONV_MODE, spatially_sparse_conv)
from warpconvnet_b200.nn.modules.base_module import BaseSpatialModule
from warpconvnet_b200.utils.ntuple import ntuple


def _parse_algo(algo, enum_cls):
    if algo is None:
        return enum_cls.TCGEN05
    if isinstance(algo, str):
        return enum_cls(algo)
    return algo  # enum member or list (lists are accepted like the reference, :130-137)


class SpatiallySparseConv(BaseSpatialModule):
    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        kernel_size: Union[int, Tuple[int, ...]],
        stride: Union[int, Tuple[int, ...]] = 1,
        dilation: Union[int, Tuple[int, ...]] = 1,
        bias: bool = True,
        transposed: bool = False,
        generative: bool = False,
        groups: int = 1,
        kernel_matmul_batch_size: int = 2,
        num_spatial_dims: Optional[int] = 3,
        fwd_algo=None,
        dgrad_algo=None,
        wgrad_algo=None,
        stride_mode: STRIDED_CONV_MODE = STRIDED_CONV_MODE.STRIDE_ONLY,
        order=None,
        compute_dtype: Optional[torch.dtype] = None,
        use_fp16_accum: Optional[bool] = None,
        implicit_matmul_fwd_block_size: Optional[int] = None,
        implicit_matmul_bwd_block_size: Optional[int] = None,
    ):
        super().__init__()
        self.num_spatial_dims = num_spatial_dims
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.groups = groups
        self.use_fp16_accum = use_fp16_accum
        if in_channels % groups != 0:
            raise ValueError(f"in_channels ({in_channels}) must be divisible by groups ({groups})")
        if out_channels % groups != 0:
            raise ValueError(f"out_channels ({out_channels}) must be divisible by groups ({groups})")
        self.kernel_size = ntuple(kernel_size, ndim=num_spatial_dims)
        self.stride = ntuple(stride, ndim=num_spatial_dims)
        self.dilation = ntuple(dilation, ndim=num_spatial_dims)
        self.transposed = transposed
        self.generative = generative
        self.kernel_matmul_batch_size = kernel_matmul_batch_size
        self.fwd_algo = _parse_algo(fwd_algo, SPARSE_CONV_AB_ALGO_MODE)
        self.dgrad_algo = _parse_algo(dgrad_algo, SPARSE_CONV_AB_ALGO_MODE)
        self.wgrad_algo = _parse_algo(wgrad_algo, SPARSE_CONV_ATB_ALGO_MODE)
        self.stride_mode = stride_mode
        self.order = order
        self.compute_dtype = compute_dtype
        self.implicit_matmul_fwd_block_size = implicit_matmul_fwd_block_size
        self.implicit_matmul_bwd_block_size = implicit_matmul_bwd_block_size

        K = int(np.prod(self.kernel_size))
        if groups == 1:
            self.weight = nn.Parameter(torch.randn(K, in_channels, out_channels))
        else:
            self.weight = nn.Parameter(
                torch.randn(K, groups, in_channels // groups, out_channels // groups))
        self.bias: Optional[nn.Parameter] = nn.Parameter(torch.randn(out_channels)) if bias else None
        self.reset_parameters()

    def __repr__(self):
        s = (f"{self.__class__.__name__}(in_channels={self.in_channels}, "
             f"out_channels={self.out_channels}, kernel_size={self.kernel_size}")
        if any(v != 1 for v in self.stride):
            s += f", stride={self.stride}"
        if any(v != 1 for v in self.dilation):
            s += f", dilation={self.dilation}"
        if self.groups != 1:
            s += f", groups={self.groups}"
        if self.transposed:
            s += f", transposed={self.transposed}"
        if self.generative:
            s += f", generative={self.generative}"
        return s + ")"

    def _calculate_fan_in_and_fan_out(self):
        rf = int(np.prod(self.kernel_size))
        return (self.in_channels // self.groups) * rf, (self.out_channels // self.groups) * rf

    def _custom_kaiming_uniform_(self, tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
        fan_in, fan_out = self._calculate_fan_in_and_fan_out()
        fan = fan_in if mode == "fan_in" else fan_out
        std = calculate_gain(nonlinearity, a) / math.sqrt(fan)
        bound = math.sqrt(self.num_spatial_dims) * std
        with torch.no_grad():
            return tensor.uniform_(-bound, bound)

    @torch.no_grad()
    def reset_parameters(self):
        self._custom_kaiming_uniform_(self.weight, a=math.sqrt(5),
                                      mode="fan_out" if self.transposed else "fan_in")
        if self.bias is not None:
            fan_in, _ = self._calculate_fan_in_and_fan_out()
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    def forward(self, input_sparse_tensor: Voxels,
                output_spatially_sparse_tensor: Optional[Voxels] = None):
        return spatially_sparse_conv(
            input_sparse_tensor=input_sparse_tensor,
            weight=self.weight,
            kernel_size=self.kernel_size,
            stride=self.stride,
            kernel_dilation=self.dilation,
            bias=self.bias,
            groups=self.groups,
            kernel_matmul_batch_size=self.kernel_matmul_batch_size,
            output_spatially_sparse_tensor=output_spatially_sparse_tensor,
            transposed=self.transposed,
            generative=self.generative,
            fwd_algo=self.fwd_algo,
            dgrad_algo=self.dgrad_algo,
            wgrad_algo=self.wgrad_algo,
            stride_mode=self.stride_mode,
            order=self.order,
            compute_dtype=self.compute_dtype,
            use_fp16_accum=self.use_fp16_accum,
            implicit_matmul_fwd_block_size=self.implicit_matmul_fwd_block_size,
            implicit_matmul_bwd_block_size=self.implicit_matmul_bwd_block_size,
        )


class SparseConv2d(SpatiallySparseConv):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, bias=True,
                 transposed=False, generative: bool = False, groups: int = 1,
                 stride_mode: STRIDED_CONV_MODE = STRIDED_CONV_MODE.STRIDE_ONLY, fwd_algo=None,
                 dgrad_algo=None, wgrad_algo=None, kernel_matmul_batch_size: int = 2, order=None,
                 compute_dtype: Optional[torch.dtype] = None,
                 use_fp16_accum: Optional[bool] = None):
        super().__init__(in_channels=in_channels, out_channels=out_channels,
                         kernel_size=kernel_size, stride=stride, dilation=dilation, bias=bias,
                         transposed=transposed, generative=generative, groups=groups,
                         num_spatial_dims=2, stride_mode=stride_mode, fwd_algo=fwd_algo,
                         dgrad_algo=dgrad_algo, wgrad_algo=wgrad_algo,
                         kernel_matmul_batch_size=kernel_matmul_batch_size, order=order,
                         compute_dtype=compute_dtype, use_fp16_accum=use_fp16_accum)


class SparseConv3d(SpatiallySparseConv):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, bias=True,
                 transposed=False, generative: bool = False, groups: int = 1,
                 stride_mode: STRIDED_CONV_MODE = STRIDED_CONV_MODE.STRIDE_ONLY, fwd_algo=None,
                 dgrad_algo=None, wgrad_algo=None, kernel_matmul_batch_size: int = 2, order=None,
                 compute_dtype: Optional[torch.dtype] = None,
                 use_fp16_accum: Optional[bool] = None):
        super().__init__(in_channels=in_channels, out_channels=out_channels,
                         kernel_size=kernel_size, stride=stride, dilation=dilation, bias=bias,
                         transposed=transposed, generative=generative, groups=groups,
                         num_spatial_dims=3, stride_mode=stride_mode, fwd_algo=fwd_algo,
                         dgrad_algo=dgrad_algo, wgrad_algo=wgrad_algo,
                         kernel_matmul_batch_size=kernel_matmul_batch_size, order=order,
                         compute_dtype=compute_dtype, use_fp16_accum=use_fp16_accum)
