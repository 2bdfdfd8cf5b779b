# SPDX-License-Identifier: Apache-2.0
from warpconvnet_b200.geometry.base.batched import CatFeatures, Features, to_batched_features  # noqa: F401
