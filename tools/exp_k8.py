# SPDX-License-Identifier: Apache-2.0
"""Per-role wait counters of the gather-GEMM kernel on the 2^3 stride-2 layers of MinkUNet-14
(single-step tiles: one offset per output row). Bring-up only: needs WCN_KERNEL_COUNTERS=1 build."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from minkunet14 import surface_scene  # noqa: E402
from warpconvnet_b200.geometry.types.voxels import Voxels  # noqa: E402
from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d  # noqa: E402

scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
coords = [surface_scene(548, s).cuda() for s in range(scenes)]
names = ["prod_total", "prod_wait_empty", "mma_total", "mma_wait_full", "mma_wait_accempty",
         "epi_total", "epi_wait_accfull", "prod_ns"]
dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
os.environ["WCN_DEBUG_PTR"] = str(dbg.data_ptr())
os.environ["WCN_DEBUG"] = "0"
for c in (32, 96):
    feats = [torch.randn(len(x), c, device="cuda").bfloat16() for x in coords]
    fine = Voxels(coords, feats)
    down = SparseConv3d(c, c, 2, 2, bias=False).cuda().bfloat16()
    up = SparseConv3d(c, c, 2, 2, transposed=True, bias=False).cuda().bfloat16()
    for stats in (False, True):
        down.emit_bn_stats = up.emit_bn_stats = stats
        with torch.no_grad():
            for name, fn in (("down (fine -> coarse)", lambda: down(fine)),):
                coarse = fn()
            for name, fn in (("down (fine -> coarse)", lambda: down(fine)),
                             ("up   (coarse -> fine)", lambda: up(coarse, fine))):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                dbg.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda._sleep(400000)
                a.record(); out = fn(); b.record()
                torch.cuda.synchronize()
                d = dbg.view(148, 16).cpu().numpy()
                print(f"c={c} stats={stats} {name}: rows {fine.feature_tensor.shape[0]} / {coarse.feature_tensor.shape[0]}"
                      f"  whole call {a.elapsed_time(b) * 1e3:.1f} us")
                for i, nm in enumerate(names):
                    print(f"      {nm:18s} mean={d[:, i].mean():10.0f} min={d[:, i].min():10d} max={d[:, i].max():10d}")
