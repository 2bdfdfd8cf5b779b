# SPDX-License-Identifier: Apache-2.0
"""Per-kernel CUDA time of the MinkUNet-14 (config C4) step through torch.profiler (CUPTI, no
replay): where the 32 ms go. Prints a table sorted by total time.  python tools/profile_c4.py [scenes]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from minkunet14 import MinkUNet14, surface_scene  # noqa: E402
from warpconvnet_b200.geometry.types.voxels import Voxels  # noqa: E402

scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
coords = [surface_scene(548, s).cuda() for s in range(scenes)]
feats = [torch.randn(len(c), 3, device="cuda") for c in coords]
net = MinkUNet14(3, 20).cuda()
opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)


def step():
    x = Voxels(coords, feats)
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(x)
    out.feature_tensor.float().square().mean().backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages():
    t = getattr(ev, "device_time_total", 0) or getattr(ev, "cuda_time_total", 0)
    if t > 0 and ev.device_type.name == "CUDA":
        rows.append((t / N, ev.count // N, ev.key))
rows.sort(reverse=True)
total = sum(r[0] for r in rows)
print(f"total kernel time per step: {total / 1e3:.2f} ms over {sum(r[1] for r in rows)} launches")
for t, cnt, name in rows[:45]:
    print(f"{t / 1e3:8.3f} ms {100 * t / total:5.1f}%  x{cnt:4d}  {name[:110]}")
