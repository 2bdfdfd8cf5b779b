# SPDX-License-Identifier: Apache-2.0
import torch.nn as nn


class BaseSpatialModule(nn.Module):
    """Base module for spatial features (warpconvnet/nn/modules/base_module.py:12-22)."""

    @property
    def device(self):
        return next(self.parameters()).device
