# SPDX-License-Identifier: Apache-2.0
"""Radius search / kNN timing at BASELINE config C5 scale (bring-up only)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from warpconvnet_b200 import _ops  # noqa: E402

for n in (125000, 1000000):
    pts = torch.rand(n, 3, device="cuda")
    off = torch.tensor([0, n])
    for r in (0.01, 0.02):
        for _ in range(2):
            _ops.radius_search(pts, off, pts, off, r)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            idx, dist, splits = _ops.radius_search(pts, off, pts, off, r)
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / 5
        print(f"n={n} radius={r}: {t * 1e3:7.2f} ms  pairs={idx.numel()}  "
              f"({idx.numel() / n:.1f} per query)  {n / t / 1e6:.1f} M queries/s")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        _ops.knn_search(pts, off, pts, off, 16)
    torch.cuda.synchronize()
    print(f"n={n} knn k=16: {(time.perf_counter() - t0) / 5 * 1e3:7.2f} ms")
