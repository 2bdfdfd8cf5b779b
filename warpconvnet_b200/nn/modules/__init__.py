# SPDX-License-Identifier: Apache-2.0
from .sparse_conv import SparseConv2d, SparseConv3d, SpatiallySparseConv  # noqa: F401
