# SPDX-License-Identifier: Apache-2.0
"""Bandwidth experiments on the forward gather-GEMM (bring-up only; the WCN_DEBUG / WCN_STAGES
switches need a library built with `WCN_BRINGUP=1 warpconvnet_b200/csrc/build.sh` — the default
build ignores them and this script then times the default configuration only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import random_coords, surface_coords  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map  # noqa: E402


def morton(c):
    def part(x):
        x = x.astype(np.uint64) & np.uint64(0x1FFFFF)
        for sh, m in ((32, 0x1F00000000FFFF), (16, 0x1F0000FF0000FF), (8, 0x100F00F00F00F00F),
                      (4, 0x10C30C30C30C30C3), (2, 0x1249249249249249)):
            x = (x | (x << np.uint64(sh))) & np.uint64(m)
        return x
    return part(c[:, 0]) | (part(c[:, 1]) << np.uint64(1)) | (part(c[:, 2]) << np.uint64(2))


_FLUSH = None
_HEAT = None


def timeit(fn, iters=20):
    """median CUDA-event time (us) with warmed clocks and an L2 flush before every call"""
    global _FLUSH, _HEAT
    if _FLUSH is None:
        _FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        _HEAT = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    # NOTE: no dense-matmul "heating": sustained tensor load drives this part into its power cap
    # (1.9 GHz -> 1.2 GHz, tools/exp_clock.py) and it needs > 50 ms to recover.
    for _ in range(5):
        fn()
    evs = []
    for _ in range(iters):
        _FLUSH.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e3


def main():
    cin = cout = 128
    for name, c in (("S", surface_coords(448, 0)), ("R", random_coords(200000, 0.3, 0))):
        n = len(c)
        bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), c], 1)).cuda()
        km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
        L = int(km.offsets[-1])
        x = torch.randn(n, cin, device="cuda").bfloat16()
        w = (torch.randn(27, 1, cin, cout, device="cuda") * 0.02).bfloat16()
        img = _ops.weight_image(w, 27, 1, cin, cout, False)
        table = km.pair_table(n)
        plans = {"mask256": _ops.build_tile_plan(table, tile_rows=256),
                 "mask128": _ops.build_tile_plan(table, tile_rows=128)}
        # reference result (explicit torch on the GPU) to make sure the variants stay correct
        ref = torch.zeros(n, cout, device="cuda")
        for k in range(27):
            s0, s1 = int(km.offsets[k]), int(km.offsets[k + 1])
            ref.index_add_(0, km.out_maps[s0:s1].long(),
                           x[km.in_maps[s0:s1].long()].float() @ w[k, 0].float())
        for pname, plan in plans.items():
            nk = plan.tile_nk.sum().item()
            print(f"[{name}] plan={pname}: tiles={plan.num_tiles} steps={nk} "
                  f"waste={nk * plan.tile_rows / L:.3f}")
            for dbg, stages in ((0, 0),):
                os.environ["WCN_DEBUG"] = str(dbg)
                if stages:
                    os.environ["WCN_STAGES"] = str(stages)
                else:
                    os.environ.pop("WCN_STAGES", None)
                y = _ops.gather_gemm(x, img, plan, 1, cin, cout)
                err = float((y.float() - ref).abs().max() / ref.abs().max())
                t = timeit(lambda: _ops.gather_gemm(x, img, plan, 1, cin, cout))
                print(f"   dbg={dbg:2d} stages={stages or 'max'}: {t:8.1f} us  "
                      f"({2 * L * cin * cout / t / 1e6:7.1f} TFLOP/s algorithmic) relerr={err:.1e}")
    os.environ.pop("WCN_DEBUG", None)


if __name__ == "__main__":
    main()
