# SPDX-License-Identifier: Apache-2.0
from .sparse_conv import SparseConv2d, SparseConv3d, SpatiallySparseConv  # noqa: F401
from .mlp import MLPBlock  # noqa: F401
from .point_conv import PointConv  # noqa: F401
from .normalizations import BatchNorm  # noqa: F401
from .activations import ReLU  # noqa: F401
from .sparse_conv_depth import (SparseDepthwiseConv2d, SparseDepthwiseConv3d,  # noqa: F401
                                SpatiallySparseDepthwiseConv)
from .sparse_pool import SparseAvgPool, SparseMaxPool, SparseMinPool, SparsePool  # noqa: F401
