# SPDX-License-Identifier: Apache-2.0
"""Eager (non-graph) step time of the C3 hot path under launch-configuration A/Bs (bring-up):
WCN_PDL_OFF=1 (bring-up build) drops the programmatic-launch attribute; argv[1] = 0/1 puts the
statistics pass of submanifold maps on the main / side stream.  python tools/exp_eager.py 1"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CIN, COUT, K, KS, make_coords, make_tensors  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
import warpconvnet_b200.geometry.coords.search.torch_discrete as td  # noqa: E402
from warpconvnet_b200.nn.functional.sparse_conv import sparse_conv_wgrad  # noqa: E402

td._STATS_ON_SIDE = (sys.argv[1] == "1") if len(sys.argv) > 1 else True
c = make_coords("S", 0)
n = len(c)
bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), c], 1)).cuda()
xh, wh, gh = make_tensors(n, 0)
x, w, gy = xh.cuda().bfloat16(), wh.cuda().bfloat16(), gh.cuda().bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def step():
    km = td.generate_kernel_map(bc, bc, (1, 1, 1), (KS,) * 3, same_coords=True)
    plan = km.fwd_plan(n)
    img, img_t = _ops.weight_image_pair(w.view(K, 1, CIN, COUT), K, 1, CIN, COUT, w.dtype)
    _ops.gather_gemm(x, img, plan, 1, CIN, COUT)
    sparse_conv_wgrad(x, gy, (K, CIN, COUT), km)
    bplan, kflip = km.bwd_plan(n)
    _ops.gather_gemm(gy, img_t, bplan, 1, COUT, CIN, kflip=kflip)


for _ in range(5):
    step()
torch.cuda.synchronize()
for flushed in (True, False):
    ts = []
    for _ in range(30):
        if flushed:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"stats_on_side={td._STATS_ON_SIDE} pdl_off={os.environ.get('WCN_PDL_OFF', '0')} "
          f"flush={flushed}: median {np.median(ts):.3f} ms  min {np.min(ts):.3f}  max {np.max(ts):.3f}", flush=True)

# bench-style loop: no synchronisation between steps (the host runs ahead), events per step
for reps in range(3):
    evs = []
    for _ in range(20):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    print(f"  back-to-back x20: mean {np.mean(ts):.3f} ms  median {np.median(ts):.3f}  max {np.max(ts):.3f}", flush=True)
