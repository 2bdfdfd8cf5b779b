# SPDX-License-Identifier: Apache-2.0
"""Host-side (Python) profile of one MinkUNet-14 step (bring-up only)."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from minkunet14 import MinkUNet14, surface_scene  # noqa: E402
from warpconvnet_b200.geometry.types.voxels import Voxels  # noqa: E402

scenes, extent = (int(sys.argv[1]) if len(sys.argv) > 1 else 8), 548
coords = [surface_scene(extent, s).cuda() for s in range(scenes)]
feats = [torch.randn(len(c), 3, device="cuda") for c in coords]
net = MinkUNet14(3, 20).cuda()
opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)


def step():
    x = Voxels(coords, feats)
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(x)
    loss = out.feature_tensor.float().square().mean()
    loss.backward()
    opt.step()
    return loss


for _ in range(4):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
st.sort_stats("cumtime").print_stats(40)
