# SPDX-License-Identifier: Apache-2.0
"""``ReLU`` on Geometry features (warpconvnet/nn/modules/activations.py:36-53). Prefer
``BatchNorm(relu=True)`` / the conv epilogue's fused ReLU where a norm or conv precedes it."""
import torch.nn as nn

from warpconvnet_b200.geometry.base.geometry import Geometry
from warpconvnet_b200.nn.modules.base_module import BaseSpatialModule


class ReLU(BaseSpatialModule):
    def __init__(self, inplace: bool = False):
        super().__init__()
        self.relu = nn.ReLU(inplace=inplace)

    def __repr__(self):
        return f"{self.__class__.__name__}(inplace={self.relu.inplace})"

    def forward(self, input):
        if isinstance(input, Geometry):
            return input.replace(batched_features=self.relu(input.feature_tensor))
        return self.relu(input)
