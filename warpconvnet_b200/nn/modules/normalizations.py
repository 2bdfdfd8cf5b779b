# SPDX-License-Identifier: Apache-2.0
"""``BatchNorm`` for Geometry features (drop-in for warpconvnet/nn/modules/normalizations.py:30-67:
same constructor, same ``norm.*`` parameter / buffer names so state dicts interchange) computed by
the fused row-streaming kernels, plus the fused tails the reference's MinkUNet blocks spell as
separate modules (``relu=True``, ``residual=``)."""
from typing import Optional, Union

import torch
import torch.nn as nn
from torch import Tensor

from warpconvnet_b200.geometry.base.geometry import Geometry
from warpconvnet_b200.nn.functional.normalizations import batch_norm_act
from warpconvnet_b200.nn.modules.base_module import BaseSpatialModule


class NormalizationBase(BaseSpatialModule):
    def __init__(self, norm: nn.Module):
        super().__init__()
        self.norm = norm

    def __repr__(self):
        return f"{self.__class__.__name__}({self.norm})"


class BatchNorm(NormalizationBase):
    """``nn.BatchNorm1d`` semantics on the feature matrix of a Geometry (or a plain tensor).

    ``relu=True`` fuses the activation that follows it in a ConvBlock; ``forward(x, residual=r)``
    fuses the identity add of a BasicBlock (``relu(bn(x) + r)``)."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1,
                 relu: bool = False):
        super().__init__(nn.BatchNorm1d(num_features, eps=eps, momentum=momentum))
        self.relu = relu

    def forward(self, input: Union[Geometry, Tensor],
                residual: Optional[Union[Geometry, Tensor]] = None):
        bn = self.norm
        feats = input.feature_tensor if isinstance(input, Geometry) else input
        res = residual.feature_tensor if isinstance(residual, Geometry) else residual
        momentum = bn.momentum
        if self.training and bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
            if momentum is None:  # cumulative moving average, as nn.BatchNorm1d
                momentum = 1.0 / float(bn.num_batches_tracked)
        # statistics left behind by the producing conv's epilogue (SparseConv3d.emit_bn_stats):
        # valid for exactly this tensor in exactly this state
        pre = getattr(feats, "_wcn_bn_sums", None)
        sums = pre[0] if (pre is not None and pre[1] == feats.shape[0]
                          and pre[2] == feats._version) else None
        out = batch_norm_act(feats, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                             training=self.training or bn.running_mean is None,
                             momentum=0.0 if momentum is None else momentum, eps=bn.eps,
                             relu=self.relu, residual=res, sums=sums)
        if isinstance(input, Geometry):
            return input.replace(batched_features=out)
        return out
