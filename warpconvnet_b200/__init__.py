# SPDX-License-Identifier: Apache-2.0
"""warpconvnet_b200 — B200-native (sm_100a) sparse-convolution hot path with the user-facing API of
NVlabs/WarpConvNet's ``SparseConv3d`` / ``PointConv`` / ``Voxels`` / ``Points``.

Importing this package loads ``csrc/libwcn_b200.so``; there is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the native library is missing)
from ._lib import version as native_version  # noqa: F401

__version__ = "0.1.0"
