# SPDX-License-Identifier: Apache-2.0
from warpconvnet_b200.geometry.coords.integer import RealCoords  # noqa: F401
