# SPDX-License-Identifier: Apache-2.0
"""Depthwise conv kernels: time and achieved gather bandwidth (bring-up only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import surface_coords  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
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


c = surface_coords(448, 0)
n = len(c)
bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), c], 1)).cuda()
km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
L = int(km.offsets[-1])
table = km.pair_table(n)
for C in (32, 64, 128, 256):
    x = torch.randn(n, C, device="cuda").bfloat16()
    gy = torch.randn(n, C, device="cuda").bfloat16()
    w = torch.randn(27, C, device="cuda") * 0.2
    plan = km.fwd_plan(n)
    t_p = timed(lambda: _ops.depthwise_conv_plan(x, w, plan))
    t_f = timed(lambda: _ops.depthwise_conv(x, w, table))
    t_d = timed(lambda: _ops.depthwise_conv(gy, w, table, kflip=True))
    t_w = timed(lambda: _ops.depthwise_wgrad(x, gy, table))
    t_wp = timed(lambda: _ops.depthwise_wgrad_plan(x, gy, plan))
    gb = L * C * 2
    alg = n * C * 2 * 2 + 27 * n * 4
    print(f"C={C:4d}: plan fwd {t_p * 1e6:7.1f} us ({gb / t_p / 1e12:4.2f} TB/s gathered) | table fwd {t_f * 1e6:7.1f} us ({gb / t_f / 1e12:4.2f} TB/s gathered, "
          f"{alg / t_f / 1e12:4.2f} TB/s algorithmic)  dgrad {t_d * 1e6:7.1f} us  "
          f"wgrad {t_w * 1e6:7.1f} us ({gb / t_w / 1e12:4.2f} TB/s gathered)  plan wgrad {t_wp * 1e6:7.1f} us")
