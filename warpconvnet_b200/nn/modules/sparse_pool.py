# SPDX-License-Identifier: Apache-2.0
"""Pooling modules over ``sparse_reduce`` (drop-in for warpconvnet/nn/modules/sparse_pool.py:20-92)."""
from warpconvnet_b200.geometry.types.voxels import Voxels
from warpconvnet_b200.nn.functional.sparse_pool import sparse_reduce
from warpconvnet_b200.nn.modules.base_module import BaseSpatialModule


class SparsePool(BaseSpatialModule):
    """Reduce the features of a ``Voxels`` over ``kernel_size`` windows placed every ``stride``."""

    def __init__(self, kernel_size, stride, reduce: str = "max"):
        super().__init__()
        self.kernel_size = kernel_size
        self.stride = stride
        self.reduce = reduce

    def __repr__(self):
        return (f"{self.__class__.__name__}(kernel_size={self.kernel_size}, stride={self.stride}, "
                f"reduce={self.reduce})")

    def forward(self, st: Voxels) -> Voxels:
        return sparse_reduce(st, self.kernel_size, self.stride, self.reduce)


class SparseMaxPool(SparsePool):
    def __init__(self, kernel_size, stride):
        super().__init__(kernel_size, stride, "max")


class SparseMinPool(SparsePool):
    def __init__(self, kernel_size, stride):
        super().__init__(kernel_size, stride, "min")


class SparseAvgPool(SparsePool):
    def __init__(self, kernel_size, stride):
        super().__init__(kernel_size, stride, "mean")
