# SPDX-License-Identifier: Apache-2.0
from .detail.unified import (  # noqa: F401
    SPARSE_CONV_AB_ALGO_MODE,
    SPARSE_CONV_ATB_ALGO_MODE,
    UnifiedSpatiallySparseConvFunction,
    sparse_conv_dgrad,
    sparse_conv_forward,
    sparse_conv_wgrad,
)
from .helper import (  # noqa: F401
    STRIDED_CONV_MODE,
    generate_output_coords_and_kernel_map,
    spatially_sparse_conv,
)
