# SPDX-License-Identifier: Apache-2.0
"""wgrad kernel variants (bring-up only; WCN_DEBUG flags 64 / 128, honoured only by a library
built with `WCN_BRINGUP=1 warpconvnet_b200/csrc/build.sh`)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import random_coords, surface_coords  # noqa: E402
from exp_fwd import timeit  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map  # noqa: E402

cin = cout = 128
for name, c in (("S", surface_coords(448, 0)), ("R", random_coords(200000, 0.3, 0))):
    n = len(c)
    bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), c], 1)).cuda()
    km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
    L = int(km.offsets[-1])
    x = torch.randn(n, cin, device="cuda").bfloat16()
    gy = torch.randn(n, cout, device="cuda").bfloat16()
    ref = torch.zeros(27, cin, cout, device="cuda")
    for k in range(27):
        s0, s1 = int(km.offsets[k]), int(km.offsets[k + 1])
        ref[k] = x[km.in_maps[s0:s1].long()].float().T @ gy[km.out_maps[s0:s1].long()].float()
    dw = torch.zeros(27, 1, cin, cout, device="cuda")
    for dbg in (0,):
        os.environ["WCN_DEBUG"] = str(dbg)
        dw.zero_()
        _ops.wgrad(x, gy, km.in_maps, km.out_maps, km.offsets_dev, 27, 1, cin, cout, dw=dw)
        err = float((dw.view(27, cin, cout) - ref).abs().max() / ref.abs().max())
        t = timeit(lambda: _ops.wgrad(x, gy, km.in_maps, km.out_maps, km.offsets_dev, 27, 1, cin,
                                      cout, dw=dw))
        print(f"[{name}] wgrad dbg={dbg:3d}: {t:8.1f} us ({2 * L * cin * cout / t / 1e6:7.1f} TFLOP/s) "
              f"relerr={err:.1e}")
    dbg_buf = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    os.environ["WCN_DEBUG_PTR"] = str(dbg_buf.data_ptr())
    names = ["prod_total", "prod_wait_empty", "mma_total", "mma_wait_full", "mma_wait_accempty",
             "epi_total", "epi_wait_accfull", "stages"]
    for dbg in (1024, 1024 + 768):
        os.environ["WCN_DEBUG"] = str(dbg)
        dbg_buf.zero_()
        _ops.wgrad(x, gy, km.in_maps, km.out_maps, km.offsets_dev, 27, 1, cin, cout, dw=dw)
        torch.cuda.synchronize()
        d = dbg_buf.view(148, 8).cpu().numpy()
        print(f"[{name}] counters dbg={dbg}")
        for i, nm in enumerate(names):
            print(f"    {nm:18s} mean={d[:, i].mean():10.0f} min={d[:, i].min():10d} max={d[:, i].max():10d}")
    os.environ.pop("WCN_DEBUG_PTR", None)
os.environ.pop("WCN_DEBUG", None)
