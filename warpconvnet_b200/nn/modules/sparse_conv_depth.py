# SPDX-License-Identifier: Apache-2.0
"""Depthwise sparse convolution modules.

API-compatible with the reference's ``SpatiallySparseDepthwiseConv`` / ``SparseDepthwiseConv2d`` /
``SparseDepthwiseConv3d`` (warpconvnet/nn/modules/sparse_conv_depth.py:34-336): identical
constructor keywords, a ``weight`` of shape [K, C], an optional ``bias`` of shape [C] and the same
uniform initialisation bounds, so checkpoints move both ways. The arithmetic runs on the gather-FMA
kernels of ``csrc/conv_depthwise.cu`` over the kernel map's mask-sorted tile plan.
"""
import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from warpconvnet_b200.geometry.coords.integer import IntCoords
from warpconvnet_b200.geometry.types.voxels import Voxels
from warpconvnet_b200.nn.functional.sparse_conv import helper as _helper
from warpconvnet_b200.nn.functional.sparse_conv_depth import (
    SPARSE_DEPTHWISE_CONV_BWD_ALGO_MODE as _BwdMode,
    SPARSE_DEPTHWISE_CONV_FWD_ALGO_MODE as _FwdMode,
    spatially_sparse_depthwise_conv)
from warpconvnet_b200.nn.modules.base_module import BaseSpatialModule
from warpconvnet_b200.utils.ntuple import ntuple

STRIDED_CONV_MODE = _helper.STRIDED_CONV_MODE


def _uniform_bound(kernel_volume: int, ndim: int, slope: float = math.sqrt(5)) -> float:
    # kaiming-uniform with a leaky-relu gain, fan = one kernel per channel, scaled by sqrt(ndim)
    gain = nn.init.calculate_gain("leaky_relu", slope)
    return math.sqrt(ndim) * gain / math.sqrt(kernel_volume)


class SpatiallySparseDepthwiseConv(BaseSpatialModule):
    """y[out, c] = bias[c] + sum_k x[in_k(out), c] * weight[k, c] on ``Voxels``."""

    def __init__(self, channels, kernel_size, stride=1, dilation=1, bias=True, transposed=False,
                 generative=False, num_spatial_dims=3, fwd_algo=None, bwd_algo=None,
                 stride_mode=STRIDED_CONV_MODE.STRIDE_ONLY, stride_reduce="max", order=None,
                 compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        nd = int(num_spatial_dims)
        geometry = {"kernel_size": kernel_size, "stride": stride, "dilation": dilation}
        for key, value in geometry.items():
            setattr(self, key, ntuple(value, ndim=nd))
        flags = dict(num_spatial_dims=nd, channels=channels, in_channels=channels,
                     out_channels=channels, transposed=bool(transposed),
                     generative=bool(generative), stride_reduce=stride_reduce,
                     stride_mode=stride_mode, order=order, compute_dtype=compute_dtype,
                     # single backend: algorithm selectors are kept only for call-site parity
                     fwd_algo=_FwdMode.AUTO if fwd_algo is None else fwd_algo,
                     bwd_algo=_BwdMode.AUTO if bwd_algo is None else bwd_algo)
        for key, value in flags.items():
            setattr(self, key, value)
        volume = int(np.prod(self.kernel_size))
        self.weight = nn.Parameter(torch.empty(volume, channels))
        self.bias = nn.Parameter(torch.empty(channels)) if bias else None
        self.reset_parameters()

    # -- initialisation ---------------------------------------------------------------------
    def _calculate_fan_in_and_fan_out(self):
        volume = int(np.prod(self.kernel_size))
        return volume, volume

    @torch.no_grad()
    def reset_parameters(self):
        volume, _ = self._calculate_fan_in_and_fan_out()
        limit = _uniform_bound(volume, self.num_spatial_dims)
        self.weight.uniform_(-limit, limit)
        if self.bias is not None:
            b = 1.0 / math.sqrt(volume) if volume > 0 else 0.0
            self.bias.uniform_(-b, b)

    def __repr__(self):
        ones = (1,) * self.num_spatial_dims
        parts = [f"channels={self.channels}", f"kernel_size={self.kernel_size}"]
        optional = (("stride", self.stride, ones), ("dilation", self.dilation, ones),
                    ("transposed", self.transposed, False), ("generative", self.generative, False))
        parts += [f"{name}={val}" for name, val, default in optional if val != default]
        if self.bias is None:
            parts.append("bias=False")
        return f"{type(self).__name__}({', '.join(parts)})"

    # -- forward ----------------------------------------------------------------------------
    def _output_stride(self, x: Voxels, target: Optional[Voxels]):
        base = x.tensor_stride or (1,) * self.num_spatial_dims
        if not self.transposed:
            return tuple(a * b for a, b in zip(self.stride, base))
        if target is not None and target.tensor_stride is not None:
            return target.tensor_stride
        return (1,) * self.num_spatial_dims

    def forward(self, input_sparse_tensor: Voxels,
                output_spatially_sparse_tensor: Optional[Voxels] = None) -> Voxels:
        x, target = input_sparse_tensor, output_spatially_sparse_tensor
        out_bc, out_offsets, kmap = _helper.generate_output_coords_and_kernel_map(
            input_sparse_tensor=x, kernel_size=self.kernel_size, kernel_dilation=self.dilation,
            stride=self.stride, generative=self.generative, transposed=self.transposed,
            output_spatially_sparse_tensor=target, stride_mode=self.stride_mode, order=self.order)
        feats = spatially_sparse_depthwise_conv(x.feature_tensor, self.weight, kmap,
                                                out_bc.shape[0], compute_dtype=self.compute_dtype)
        if self.bias is not None:
            feats = feats + self.bias.to(feats.dtype)
        if out_bc is x.batch_indexed_coordinates:      # submanifold: keep the coordinate object
            coords = x.batched_coordinates
        else:
            host_offsets = out_offsets if out_offsets.device.type == "cpu" else out_offsets.cpu()
            coords = IntCoords(out_bc[:, 1:].contiguous(), offsets=host_offsets)
            coords._bcoords = out_bc
        return x.replace(batched_coordinates=coords, batched_features=feats,
                         tensor_stride=self._output_stride(x, target))


def _fixed_dims(nd: int, name: str):
    def __init__(self, channels, kernel_size, stride=1, dilation=1, bias=True, transposed=False,
                 generative=False, fwd_algo=None, bwd_algo=None,
                 stride_mode=STRIDED_CONV_MODE.STRIDE_ONLY, stride_reduce="max", order=None,
                 compute_dtype: Optional[torch.dtype] = None):
        SpatiallySparseDepthwiseConv.__init__(
            self, channels, kernel_size, stride=stride, dilation=dilation, bias=bias,
            transposed=transposed, generative=generative, num_spatial_dims=nd, fwd_algo=fwd_algo,
            bwd_algo=bwd_algo, stride_mode=stride_mode, stride_reduce=stride_reduce, order=order,
            compute_dtype=compute_dtype)
    return type(name, (SpatiallySparseDepthwiseConv,),
                {"__init__": __init__, "__doc__": f"{nd}-D depthwise sparse convolution.",
                 "__module__": __name__})


SparseDepthwiseConv2d = _fixed_dims(2, "SparseDepthwiseConv2d")
SparseDepthwiseConv3d = _fixed_dims(3, "SparseDepthwiseConv3d")
