# SPDX-License-Identifier: Apache-2.0
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- synthetic coordinate generators -------------------------------------------
def random_coords(n: int, occupancy: float = 0.30, seed: int = 0) -> np.ndarray:
    """(R) exactly n unique voxels drawn uniformly from a cube with the given occupancy."""
    side = int(np.ceil((n / occupancy) ** (1.0 / 3.0)))
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(side ** 3, generator=g)[:n].numpy()
    return np.stack([idx // (side * side), (idx // side) % side, idx % side], axis=1).astype(np.int32)


def surface_coords(extent: int, seed: int = 0) -> np.ndarray:
    """(S) ScanNet-like height field: z = round(12 sin(2 pi u/180 + a) + 8 cos(2 pi v/130 + b)) + 256
    over u, v in [0, extent)^2; one voxel per (u, v)."""
    rng = np.random.RandomState(seed)
    a, b = rng.uniform(0, 2 * np.pi, size=2)
    u, v = np.meshgrid(np.arange(extent), np.arange(extent), indexing="ij")
    z = np.rint(12 * np.sin(2 * np.pi * u / 180 + a) + 8 * np.cos(2 * np.pi * v / 130 + b)) + 256
    return np.stack([u.reshape(-1), v.reshape(-1), z.reshape(-1)], axis=1).astype(np.int32)


@pytest.fixture(autouse=True)
def _drain_pending_kernel_map_checks():
    """A test that provokes a deferred kernel-map error must not leak it into the next test."""
    yield
    try:
        from warpconvnet_b200.geometry.coords.search import search_results
        search_results._PENDING_STATUS.clear()
    except Exception:
        pass
