# SPDX-License-Identifier: Apache-2.0
"""Is the TMA identity path of wgrad active and what is it worth? K = 1 map: every pair is (i, i)
(bring-up only)."""
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


def timed(fn, iters=20):
    for _ in range(5):
        fn()
    evs = []
    for _ in range(iters):
        if os.environ.get("FLUSH", "0") == "1":
            FLUSH.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e3


c = surface_coords(448, 0)       # 200 k voxels, K = 1: every pair is (i, i); operands fit L2
n = len(c)
bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), c], 1)).cuda()
km = generate_kernel_map(bc, bc, (1, 1, 1), (1, 1, 1), same_coords=True)
x = torch.randn(n, 128, device="cuda").bfloat16()
gy = torch.randn(n, 128, device="cuda").bfloat16()
dw = torch.zeros(1, 1, 128, 128, device="cuda")
args = (x, gy, km._in_buf, km._out_buf, km.offsets_dev, 1, 1, 128, 128)
ref = _ops.wgrad(*args)
got = _ops.wgrad(*args, identity_k=0, status=km._hashtable.status_tensor)
print("max rel diff", float((ref - got).abs().max() / ref.abs().max()))
print("gather path  :", timed(lambda: _ops.wgrad(*args, dw=dw)), "us")
print("identity TMA :", timed(lambda: _ops.wgrad(*args, dw=dw, identity_k=0,
                                                  status=km._hashtable.status_tensor)), "us")
print("bytes: 2 x", n * 256 / 1e6, "MB")
