# SPDX-License-Identifier: Apache-2.0
"""Depthwise sparse convolution (drop-in for warpconvnet/nn/functional/sparse_conv_depth.py:227-420,
957-1010): ``y[out, c] = sum_k x[in_k(out), c] * w[k, c]`` on the kernel map's [K, M] neighbour
table, computed by the gather-FMA kernels of ``csrc/conv_depthwise.cu`` (one backend; the
reference's explicit / implicit variants and their autotuner collapse into it)."""
from enum import Enum
from typing import Optional

import torch
from torch import Tensor
from torch.autograd import Function

from warpconvnet_b200 import _ops


class SPARSE_DEPTHWISE_CONV_FWD_ALGO_MODE(Enum):
    EXPLICIT = "explicit"
    IMPLICIT = "implicit"
    AUTO = "auto"


class SPARSE_DEPTHWISE_CONV_BWD_ALGO_MODE(Enum):
    EXPLICIT = "explicit"
    IMPLICIT = "implicit"
    AUTO = "auto"


class UnifiedSpatiallySparseDepthwiseConvFunction(Function):
    @staticmethod
    def forward(ctx, in_features: Tensor, weight: Tensor, kernel_map, num_out_coords: int,
                compute_dtype: Optional[torch.dtype]):
        if not in_features.is_cuda:
            raise RuntimeError("warpconvnet_b200 runs on CUDA (sm_100a) only; no CPU fallback")
        x = in_features if compute_dtype is None else in_features.to(compute_dtype)
        x = x if x.stride(1) == 1 else x.contiguous()
        w32 = weight.detach().float().contiguous()
        # the mask-sorted tile plan is shared with the dense convs of the same resolution
        y = _ops.depthwise_conv_plan(x, w32, kernel_map.fwd_plan(num_out_coords))
        ctx.kernel_map = kernel_map
        ctx.num_in = in_features.shape[0]
        ctx.in_dtype = in_features.dtype
        ctx.save_for_backward(x, w32)
        ctx.w_dtype = weight.dtype
        return y.to(in_features.dtype)

    @staticmethod
    def backward(ctx, grad_output: Tensor):
        x, w32 = ctx.saved_tensors
        km = ctx.kernel_map
        gy = grad_output.to(x.dtype)
        gy = gy if gy.stride(1) == 1 else gy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            # (plan, kflip): for a submanifold map the reverse plan is the forward plan with the
            # offset index flipped
            plan, kflip = km.bwd_plan(ctx.num_in)
            dx = _ops.depthwise_conv_plan(gy, w32, plan, kflip=kflip)
            dx = dx.to(ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            # tile-plan walk from 64 channels up (206 vs 337 us at 128 channels on C3-S); below,
            # a thread sees too few rows per tile to amortise its shared-memory merge (184 vs 161 us
            # at 32 channels) and the dense-table kernel is used
            if x.shape[1] >= 64:
                dw = _ops.depthwise_wgrad_plan(x, gy, km.fwd_plan(gy.shape[0]))
            else:
                dw = _ops.depthwise_wgrad(x, gy, km.pair_table(gy.shape[0]))
            dw = dw.to(ctx.w_dtype)
        return dx, dw, None, None, None


def spatially_sparse_depthwise_conv(in_features: Tensor, weight: Tensor, kernel_map,
                                    num_out_coords: int, fwd_algo=None, bwd_algo=None,
                                    compute_dtype: Optional[torch.dtype] = None) -> Tensor:
    """in_features [N, C], weight [K, C] -> [M, C]; algo arguments are accepted and ignored."""
    if weight.dim() != 2 or weight.shape[1] != in_features.shape[1]:
        raise ValueError(f"depthwise weight must be [K, C={in_features.shape[1]}], "
                         f"got {tuple(weight.shape)}")
    if compute_dtype is None and torch.is_autocast_enabled():
        compute_dtype = torch.get_autocast_dtype("cuda")
    return UnifiedSpatiallySparseDepthwiseConvFunction.apply(in_features, weight, kernel_map,
                                                             num_out_coords, compute_dtype)
