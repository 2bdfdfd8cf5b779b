# SPDX-License-Identifier: Apache-2.0
"""cProfile of the public-API step (host overhead; bring-up only)."""
import cProfile, os, pstats, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import surface_coords
from warpconvnet_b200.geometry.types.voxels import Voxels
from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
c = torch.from_numpy(surface_coords(448, 0)).cuda(); n = len(c)
f = torch.randn(n, 128, device="cuda").bfloat16()
gy = torch.randn(n, 128, device="cuda").bfloat16()
conv = SparseConv3d(128, 128, 3, bias=False).cuda()
offsets = torch.tensor([0, n], dtype=torch.int64)
def step():
    ff = f.detach().requires_grad_(True)
    vox = Voxels(c, ff, offsets=offsets)
    conv.weight.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = conv(vox)
    out.feature_tensor.backward(gy)
for _ in range(5): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(50): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/50:.3f} ms/step, with drain {1e3*(t2-t0)/50:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
