# SPDX-License-Identifier: Apache-2.0
"""Kernel timeline (CUPTI through torch.profiler) of graph replays of the C3-S benchmark step:
start / duration / stream of every kernel — which launches sit on the critical path, where the
side-stream CSR branch overlaps the tile plan. python tools/trace_c3_step.py > gpurun_out/c3_timeline.txt"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map  # noqa: E402
from warpconvnet_b200.nn.functional.sparse_conv import sparse_conv_wgrad  # noqa: E402

K, CIN, COUT, KS = bench.K, bench.CIN, bench.COUT, bench.KS
dev = torch.device("cuda", 0)
coords = bench.make_coords("S", seed=0)
n = len(coords)
x_h, w_h, gy_h = bench.make_tensors(n, seed=0)
bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), coords], 1)).to(dev)
x, w, gy = x_h.to(dev).bfloat16(), w_h.to(dev).bfloat16(), gy_h.to(dev).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step():
    km = generate_kernel_map(bc, bc, (1, 1, 1), (KS,) * 3, same_coords=True)
    plan = km.fwd_plan(n)
    img, img_t = _ops.weight_image_pair(w.view(K, 1, CIN, COUT), K, 1, CIN, COUT, w.dtype)
    y = _ops.gather_gemm(x, img, plan, 1, CIN, COUT)
    dw = sparse_conv_wgrad(x, gy, (K, CIN, COUT), km)
    bplan, kflip = km.bwd_plan(n)
    dx = _ops.gather_gemm(gy, img_t, bplan, 1, COUT, CIN, kflip=kflip)
    return y, dw, dx


for _ in range(3):
    step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, capture_error_mode="relaxed"):
    out = step()
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        flush.fill_(1)
        g.replay()
    torch.cuda.synchronize()
rows = []
for ev in prof.events():
    if ev.device_type is not None and "CUDA" in str(ev.device_type):
        rows.append((ev.time_range.start, ev.time_range.end - ev.time_range.start, ev.name[:70]))
rows.sort()
# last replay only
starts = [i for i, r in enumerate(rows) if "vectorized_elementwise" in r[2] and r[1] > 40]
rows = rows[starts[-1]:]
t0 = rows[0][0] + rows[0][1]
print("start_us (after the L2 flush)  dur_us  end_us  kernel")
for st, du, nm in rows[1:]:
    print(f"{st - t0:9.1f} {du:7.1f} {st + du - t0:8.1f}  {nm}")
