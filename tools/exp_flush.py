# SPDX-License-Identifier: Apache-2.0
"""How much of the event-timed forward kernel is an artefact of what precedes it (bring-up only)?
(a) 256 MiB write flush (bench.py's method: L2 left full of DIRTY lines), (b) write flush followed by
a 256 MiB read (L2 left full of CLEAN lines), (c) no flush, back-to-back launches."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import surface_coords  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map  # noqa: E402

cin = cout = 128
c = surface_coords(448, 0)
n = len(c)
bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), c], 1)).cuda()
km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
x = torch.randn(n, cin, device="cuda").bfloat16()
w = (torch.randn(27, 1, cin, cout, device="cuda") * 0.02).bfloat16()
img = _ops.weight_image(w, 27, 1, cin, cout, False)
plan = _ops.build_tile_plan(km.pair_table(n), tile_rows=256)
y = torch.empty(n, cout, device="cuda", dtype=torch.bfloat16)
wbuf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rbuf = torch.ones(64 << 20, dtype=torch.float32, device="cuda")


def run(pre, iters=20):
    fn = lambda: _ops.gather_gemm(x, img, plan, 1, cin, cout, out=y)
    for _ in range(5):
        fn()
    evs = []
    for _ in range(iters):
        pre()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) * 1e3 for a, b in evs]
    return float(np.median(t)), float(np.min(t))


def pre_write():
    wbuf.fill_(1)


def pre_write_read():
    wbuf.fill_(1)
    rbuf.sum()


def pre_none():
    pass


def pre_prev():  # a previous launch of the same kernel keeps the GPU busy (back to back, warm L2)
    _ops.gather_gemm(x, img, plan, 1, cin, cout, out=y)


for name, pre in (("write flush", pre_write), ("write + read flush", pre_write_read),
                  ("idle GPU, warm L2", pre_none), ("back to back, warm L2", pre_prev)):
    med, mn = run(pre)
    print(f"{name:24s}: median {med:6.1f} us  min {mn:6.1f} us")
