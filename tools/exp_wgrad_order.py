# SPDX-License-Identifier: Apache-2.0
"""wgrad: offset-major vs row-block-major unit order, time with a flushed L2 (bring-up only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import random_coords, surface_coords  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map  # noqa: E402

FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=20):
    for _ in range(5):
        fn()
    evs = []
    for _ in range(iters):
        FLUSH.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e3


only = sys.argv[1] if len(sys.argv) > 1 else None
for name, c in (("S", surface_coords(448, 0)), ("R", random_coords(200000, 0.3, 0))):
    n = len(c)
    bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), c], 1)).cuda()
    km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
    x = torch.randn(n, 128, device="cuda").bfloat16()
    gy = torch.randn(n, 128, device="cuda").bfloat16()
    dw = torch.zeros(27, 1, 128, 128, device="cuda")
    args = (x, gy, km._in_buf, km._out_buf, km.offsets_dev, 27, 1, 128, 128)
    ident = {"identity_k": 13, "status": km._hashtable.status_tensor}
    variants = [("offset-major", {}), ("offset-major + identity TMA", dict(ident)),
                ("parts=2 rounds=2 + identity TMA",
                 dict(ident, row_block_prefix=km._block_prefix, row_parts=2, rounds=2))] + [
        (f"parts={p} rounds={r}", {"row_block_prefix": km._block_prefix, "row_parts": p, "rounds": r})
        for p, r in ((2, 2), (4, 4), (8, 4), (8, 8), (16, 8), (4, 8))]
    for vname, kw in variants:
        if only and only not in vname:
            continue
        t = timed(lambda: _ops.wgrad(*args, dw=dw, **kw))
        print(f"[{name}] {vname:22s} {t:7.1f} us")
