# SPDX-License-Identifier: Apache-2.0
"""Kernel-map build (+ tile plan) time vs size: achieved bytes/s against the algorithmic
236 B/voxel of DESIGN.md 4.1 (bring-up only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import surface_coords  # noqa: E402
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map  # noqa: E402

FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    evs = []
    for _ in range(iters):
        FLUSH.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e-3


use_graph = "--eager" not in sys.argv
sizes = ((448, 1), (548, 2), (548, 4), (548, 8))
if len(sys.argv) > 1 and sys.argv[1].isdigit():
    sizes = tuple(sz for sz in sizes if sz[1] == int(sys.argv[1]))
for extent, scenes in sizes:
    cs = [np.concatenate([np.full((extent * extent, 1), s, np.int32), surface_coords(extent, s)], 1)
          for s in range(scenes)]
    bc = torch.from_numpy(np.concatenate(cs)).cuda()
    n = len(bc)
    km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
    torch.cuda.synchronize()
    if use_graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="relaxed"):
            km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
        t = timed(g.replay)
    else:
        t = timed(lambda: generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True))
    L = int(km.offsets[-1])
    alg = n * (16 + 24 + 16 + 4 * 27) + 8 * L
    plan = n * (8 + 4) * 2 * 4 + 27 * n * 4 + (L + n) * 4  # sort passes + table read + step lists
    print(f"N={n:8d} L={L:9d}: map + plan {t * 1e6:8.1f} us (graph replay)  "
          f"{n / t / 1e6:7.1f} M voxels/s  {(alg + plan) / t / 1e12:5.2f} TB/s algorithmic "
          f"({(alg + plan) / 1e6:.0f} MB)")
