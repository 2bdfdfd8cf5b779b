# SPDX-License-Identifier: Apache-2.0
"""Host-side (Python) cost of one SparseConv3d fwd+bwd through the public API with the kernel map
cached: cProfile over 300 steps, top functions by cumulative time.  python tools/profile_host.py"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ref_gpu_bench import random_cube, surface  # noqa: E402
from warpconvnet_b200.geometry.types.voxels import Voxels  # noqa: E402
from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "S"
c = (surface(448, 0) if which == "S" else random_cube(200000, 88, 0)).cuda()
n = len(c)
x = torch.randn(n, 128, device="cuda").bfloat16()
gy = torch.randn(n, 128, device="cuda").bfloat16()
conv = SparseConv3d(128, 128, 3, bias=False).cuda()
vox = Voxels([c], [x])


def step():
    conv.weight.grad = None
    v = vox.replace(batched_features=vox.feature_tensor.detach().requires_grad_(True))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = conv(v)
    y.feature_tensor.backward(gy)


for _ in range(10):
    step()
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        step()
    e.record()
    host = (time.perf_counter() - t0) / 50 * 1e3
    torch.cuda.synchronize()
    print(f"{which} rep {rep}: host enqueue {host:.3f} ms/step, events {s.elapsed_time(e) / 50:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(38)
