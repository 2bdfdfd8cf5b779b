# SPDX-License-Identifier: Apache-2.0
"""``SparseDepthwiseConv3d`` (drop-in for warpconvnet/nn/modules/sparse_conv_depth.py:34-336: same
constructor, ``weight`` [K, C] / ``bias`` [C], same initialisation) on the gather-FMA kernels of
``csrc/conv_depthwise.cu``."""
import math
from typing import Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn
from torch.nn import init
from torch.nn.init import calculate_gain

from warpconvnet_b200.geometry.coords.integer import IntCoords
from warpconvnet_b200.geometry.types.voxels import Voxels
from warpconvnet_b200.nn.functional.sparse_conv.helper import (
    STRIDED_CONV_MODE, generate_output_coords_and_kernel_map)
from warpconvnet_b200.nn.functional.sparse_conv_depth import (
    SPARSE_DEPTHWISE_CONV_BWD_ALGO_MODE, SPARSE_DEPTHWISE_CONV_FWD_ALGO_MODE,
    spatially_sparse_depthwise_conv)
from warpconvnet_b200.nn.modules.base_module import BaseSpatialModule
from warpconvnet_b200.utils.ntuple import ntuple


class SpatiallySparseDepthwiseConv(BaseSpatialModule):
    def __init__(self, channels: int, kernel_size: Union[int, Tuple[int, ...]],
                 stride: Union[int, Tuple[int, ...]] = 1,
                 dilation: Union[int, Tuple[int, ...]] = 1, bias: bool = True,
                 transposed: bool = False, generative: bool = False, num_spatial_dims: int = 3,
                 fwd_algo=None, bwd_algo=None,
                 stride_mode: STRIDED_CONV_MODE = STRIDED_CONV_MODE.STRIDE_ONLY,
                 stride_reduce: str = "max", order=None,
                 compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        self.num_spatial_dims = num_spatial_dims
        self.channels = self.in_channels = self.out_channels = channels
        self.kernel_size = ntuple(kernel_size, ndim=num_spatial_dims)
        self.stride = ntuple(stride, ndim=num_spatial_dims)
        self.dilation = ntuple(dilation, ndim=num_spatial_dims)
        self.transposed = transposed
        self.generative = generative
        self.stride_reduce = stride_reduce
        # one backend: the algorithm arguments are accepted for source compatibility
        self.fwd_algo = fwd_algo if fwd_algo is not None else SPARSE_DEPTHWISE_CONV_FWD_ALGO_MODE.AUTO
        self.bwd_algo = bwd_algo if bwd_algo is not None else SPARSE_DEPTHWISE_CONV_BWD_ALGO_MODE.AUTO
        self.stride_mode = stride_mode
        self.order = order
        self.compute_dtype = compute_dtype
        kernel_volume = int(np.prod(self.kernel_size))
        self.weight = nn.Parameter(torch.randn(kernel_volume, channels))
        self.bias = nn.Parameter(torch.randn(channels)) if bias else None
        self.reset_parameters()

    def __repr__(self):
        s = f"{self.__class__.__name__}(channels={self.channels}, kernel_size={self.kernel_size}"
        if self.stride != (1,) * self.num_spatial_dims:
            s += f", stride={self.stride}"
        if self.dilation != (1,) * self.num_spatial_dims:
            s += f", dilation={self.dilation}"
        if self.transposed:
            s += f", transposed={self.transposed}"
        if self.generative:
            s += f", generative={self.generative}"
        if self.bias is None:
            s += ", bias=False"
        return s + ")"

    def _calculate_fan_in_and_fan_out(self):
        rf = int(np.prod(self.kernel_size))  # one kernel per channel
        return rf, rf

    def _custom_kaiming_uniform_(self, tensor, a=0.0, mode="fan_in", nonlinearity="leaky_relu"):
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
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            init.uniform_(self.bias, -bound, bound)

    def forward(self, input_sparse_tensor: Voxels,
                output_spatially_sparse_tensor: Optional[Voxels] = None) -> Voxels:
        bout, out_offsets, kernel_map = generate_output_coords_and_kernel_map(
            input_sparse_tensor=input_sparse_tensor, kernel_size=self.kernel_size,
            kernel_dilation=self.dilation, stride=self.stride, generative=self.generative,
            transposed=self.transposed,
            output_spatially_sparse_tensor=output_spatially_sparse_tensor,
            stride_mode=self.stride_mode, order=self.order)
        num_out = bout.shape[0]
        out = spatially_sparse_depthwise_conv(
            input_sparse_tensor.feature_tensor, self.weight, kernel_map, num_out,
            fwd_algo=self.fwd_algo, bwd_algo=self.bwd_algo, compute_dtype=self.compute_dtype)
        if self.bias is not None:
            out = out + self.bias.to(out.dtype)
        in_ts = input_sparse_tensor.tensor_stride
        if in_ts is None:
            in_ts = (1,) * self.num_spatial_dims
        if not self.transposed:
            out_ts = tuple(o * s for o, s in zip(self.stride, in_ts))
        elif (output_spatially_sparse_tensor is not None
              and output_spatially_sparse_tensor.tensor_stride is not None):
            out_ts = output_spatially_sparse_tensor.tensor_stride
        else:
            out_ts = (1,) * self.num_spatial_dims
        offs = out_offsets if out_offsets.device.type == "cpu" else out_offsets.cpu()
        if bout is input_sparse_tensor.batch_indexed_coordinates:
            coords = input_sparse_tensor.batched_coordinates
        else:
            coords = IntCoords(bout[:, 1:].contiguous(), offsets=offs)
            coords._bcoords = bout
        return input_sparse_tensor.replace(batched_coordinates=coords, batched_features=out,
                                           tensor_stride=out_ts)


class SparseDepthwiseConv2d(SpatiallySparseDepthwiseConv):
    def __init__(self, channels, kernel_size, stride=1, dilation=1, bias=True, transposed=False,
                 generative=False, fwd_algo=None, bwd_algo=None,
                 stride_mode: STRIDED_CONV_MODE = STRIDED_CONV_MODE.STRIDE_ONLY,
                 stride_reduce: str = "max", order=None,
                 compute_dtype: Optional[torch.dtype] = None):
        super().__init__(channels, kernel_size, stride, dilation, bias, transposed, generative, 2,
                         fwd_algo, bwd_algo, stride_mode, stride_reduce, order, compute_dtype)


class SparseDepthwiseConv3d(SpatiallySparseDepthwiseConv):
    def __init__(self, channels, kernel_size, stride=1, dilation=1, bias=True, transposed=False,
                 generative=False, fwd_algo=None, bwd_algo=None,
                 stride_mode: STRIDED_CONV_MODE = STRIDED_CONV_MODE.STRIDE_ONLY,
                 stride_reduce: str = "max", order=None,
                 compute_dtype: Optional[torch.dtype] = None):
        super().__init__(channels, kernel_size, stride, dilation, bias, transposed, generative, 3,
                         fwd_algo, bwd_algo, stride_mode, stride_reduce, order, compute_dtype)
