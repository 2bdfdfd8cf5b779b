# SPDX-License-Identifier: Apache-2.0
"""A/B of the wgrad unit orders on C3 (S and R): (a) offset-major pair lists, (b) unit table
(row parts x offsets, round-robin chunks) without dense rows, (c) dense-row offsets.
python tools/exp_wgrad_dense.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CIN, COUT, K, make_coords, make_tensors  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, k=20, w=5):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(k):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)) * 1e3


for dist in ("S", "R"):
    c = make_coords(dist, 0)
    n = len(c)
    bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), c], 1)).cuda()
    xh, _, gh = make_tensors(n, 0)
    x, gy = xh.cuda().bfloat16(), gh.cuda().bfloat16()
    km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
    im, om, od, bp, pt = km._in_buf, km._out_buf, km.offsets_dev, km._block_prefix, km._pair_table
    dw = torch.zeros((K, 1, CIN, COUT), dtype=torch.float32, device="cuda")
    res = {"dist": dist, "n": n}
    res["a_offset_major_us"] = timed(lambda: _ops.wgrad(x, gy, im, om, od, K, 1, CIN, COUT, dw=dw))
    for P, R in ((1, 2), (2, 2), (1, 4)):
        res[f"b_units_P{P}_R{R}_us"] = timed(lambda: _ops.wgrad(
            x, gy, im, om, od, K, 1, CIN, COUT, dw=dw, row_block_prefix=bp, row_parts=P, rounds=R))
        res[f"c_dense_P{P}_R{R}_us"] = timed(lambda: _ops.wgrad(
            x, gy, im, om, od, K, 1, CIN, COUT, dw=dw, row_block_prefix=bp, row_parts=P, rounds=R,
            pair_table=pt))
    print(json.dumps(res), flush=True)
