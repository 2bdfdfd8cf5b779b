# SPDX-License-Identifier: Apache-2.0
"""Timing of wcn_sort_rows_by_key (the mask sort of the tile plan) on C3-sized inputs (bring-up:
WCN_SORT_KEYS_PER_CTA is honoured by WCN_BRINGUP builds only).  python tools/exp_sort.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_coords  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
from warpconvnet_b200._lib import check, lib  # noqa: E402
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for dist in ("S", "R"):
    c = make_coords(dist, 0)
    n = len(c)
    bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), c], 1)).cuda()
    km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True, build_plan=False)
    keys = km._mask_keys
    rows = torch.empty(n, dtype=torch.int32, device="cuda")
    ws_bytes = lib.wcn_sort_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")

    def run():
        check(lib.wcn_sort_rows_by_key(keys.data_ptr(), n, 27, rows.data_ptr(), ws.data_ptr(), ws_bytes,
                                       _ops._stream()), "sort")
    for _ in range(5):
        run()
    ts = []
    for _ in range(20):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print(f"{dist} n={n} keys/cta={os.environ.get('WCN_SORT_KEYS_PER_CTA', 'default')}: "
          f"median {np.median(ts):.1f} us  min {np.min(ts):.1f} us", flush=True)
