# SPDX-License-Identifier: Apache-2.0
"""BASELINE config 4 timing: MinkUNet-14 shape, 8 scenes x ~300k voxels, fwd+bwd+SGD (bring-up)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from minkunet14 import MinkUNet14, surface_scene
from warpconvnet_b200.geometry.types.voxels import Voxels

scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
extent = int(sys.argv[2]) if len(sys.argv) > 2 else 548
coords = [surface_scene(extent, s).cuda() for s in range(scenes)]
feats = [torch.randn(len(c), 3, device="cuda") for c in coords]
n = sum(len(c) for c in coords)
net = MinkUNet14(3, 20).cuda()
opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)

def step():
    x = Voxels(coords, feats)          # fresh container: kernel maps are rebuilt every step
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(x)
    loss = out.feature_tensor.float().square().mean()
    loss.backward()
    opt.step()
    return loss

for i in range(4):
    t0 = time.time(); l = step(); torch.cuda.synchronize(); print(f"warmup {i}: {time.time()-t0:.3f}s loss {float(l):.4f}")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
a.record()
for _ in range(K):
    step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
print(f"C4 MinkUNet14 scenes={scenes} voxels={n}: {ms:.2f} ms/step  {n / ms / 1e3:.2f} M voxels/s  mem={torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
from torch.profiler import profile, ProfilerActivity
try:
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
except Exception as e:
    print("profiler unavailable:", e)
