# SPDX-License-Identifier: Apache-2.0
"""Achieved HBM bandwidth of the row-norm kernels (bring-up only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from warpconvnet_b200 import _ops  # noqa: E402

FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    evs = []
    for _ in range(iters):
        FLUSH.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e-3


for n, c in ((2402432, 96), (754393, 96), (754393, 32), (208065, 128)):
    x = torch.randn(n, c, device="cuda").bfloat16()
    res = torch.randn(n, c, device="cuda").bfloat16()
    dy = torch.randn(n, c, device="cuda").bfloat16()
    gamma = torch.ones(c, device="cuda")
    beta = torch.zeros(c, device="cuda")
    sums = _ops.bn_stats(x)
    scale, shift, mr = _ops.bn_finalize(sums, n, gamma, beta, 1e-5, 0.1, None, None)
    y = _ops.scale_shift_act(x, scale, shift, res, True)
    bsums = _ops.bn_bwd_reduce(dy, x, y, mr)
    B = n * c * 2
    rows = [("stats (1 read)", lambda: _ops.bn_stats(x), B),
            ("apply+res+relu (2r 1w)", lambda: _ops.scale_shift_act(x, scale, shift, res, True, out=y), 3 * B),
            ("apply+relu (1r 1w)", lambda: _ops.scale_shift_act(x, scale, shift, None, True, out=y), 2 * B),
            ("bwd reduce (3r)", lambda: _ops.bn_bwd_reduce(dy, x, y, mr), 3 * B),
            ("bwd apply (3r 1w)", lambda: _ops.bn_bwd_apply(dy, x, y, gamma, mr, bsums, True, False), 4 * B),
            ("bwd apply + dres (3r 2w)", lambda: _ops.bn_bwd_apply(dy, x, y, gamma, mr, bsums, True, True), 5 * B),
            ("bwd reduce mask-from-x (2r)", lambda: _ops.bn_bwd_reduce(dy, x, None, mr, scale, shift), 2 * B),
            ("bwd apply mask-from-x (2r 1w)", lambda: _ops.bn_bwd_apply(dy, x, None, gamma, mr, bsums, True, False, scale, shift), 3 * B),
            ("bwd reduce no mask (2r)", lambda: _ops.bn_bwd_reduce(dy, x, None, mr), 2 * B),
            ("bwd apply no mask (2r 1w)", lambda: _ops.bn_bwd_apply(dy, x, None, gamma, mr, bsums, True, False), 3 * B),
            ("bn_backward one call mask-x", lambda: _ops.bn_backward(dy, x, None, gamma, mr, scale, shift, False), 5 * B),
            ("bn_forward from sums +relu", lambda: _ops.bn_forward(x, gamma, beta, 1e-5, 0.1, None, None, None, True, sums=sums), 2 * B),
            ("torch copy (1r 1w)", lambda: y.copy_(x), 2 * B)]
    for name, fn, nbytes in rows:
        t = timed(fn)
        print(f"[{n} x {c}] {name:28s} {t * 1e6:8.1f} us  {nbytes / t / 1e12:5.2f} TB/s")
