# SPDX-License-Identifier: Apache-2.0
"""The other BASELINE configs (C2, C4, C5) on one GPU — one JSON line each. The headline bench
(bench.py) is C3; these lines feed DESIGN.md §5 and profiles/. Usage: python tools/bench_configs.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from minkunet14 import MinkUNet14, surface_scene  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
from warpconvnet_b200.geometry.coords.search.search_configs import RealSearchConfig  # noqa: E402
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map  # noqa: E402
from warpconvnet_b200.geometry.types.points import Points  # noqa: E402
from warpconvnet_b200.geometry.types.voxels import Voxels  # noqa: E402
from warpconvnet_b200.nn.functional.sparse_conv import (sparse_conv_dgrad, sparse_conv_forward,  # noqa: E402
                                                        sparse_conv_wgrad)
from warpconvnet_b200.nn.modules.point_conv import PointConv  # noqa: E402

FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, steps=10, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        if flush:
            FLUSH.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in evs]))


def c2():
    """SparseConv3d 3^3, 64 -> 128 bf16, ~100k voxels (surface 317^2): kernel map + forward."""
    c = surface_scene(317, 0).cuda()
    n = len(c)
    bc = torch.cat([torch.zeros(n, 1, dtype=torch.int32, device="cuda"), c], 1).contiguous()
    x = torch.randn(n, 64, device="cuda").bfloat16()
    w = (torch.randn(27, 64, 128, device="cuda") * (27 * 64) ** -0.5).bfloat16()
    km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
    L = int(km.offsets[-1])
    t_map = timed(lambda: generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True))
    t_fwd = timed(lambda: sparse_conv_forward(x, w, km, n))
    fl = 2.0 * L * 64 * 128
    return {"config": "C2: SparseConv3d 3^3 64->128 bf16, %d voxels (S), fwd only" % n,
            "kernel_map_plus_plan_ms": t_map, "fwd_ms": t_fwd,
            "voxels_per_s_map_plus_fwd": n / ((t_map + t_fwd) * 1e-3),
            "fwd_TFLOPs_algorithmic": fl / (t_fwd * 1e-3) / 1e12, "pairs_L": L}


def c4(scenes=8, extent=548):
    """MinkUNet-14 shape, 8 scenes x ~300k voxels, AMP bf16, fwd + bwd + SGD step."""
    coords = [surface_scene(extent, s).cuda() for s in range(scenes)]
    feats = [torch.randn(len(c), 3, device="cuda") for c in coords]
    n = sum(len(c) for c in coords)
    net = MinkUNet14(3, 20).cuda()
    opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)

    def step():
        x = Voxels(coords, feats)  # fresh container: every kernel map is rebuilt each step
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(x)
        out.feature_tensor.float().square().mean().backward()
        opt.step()

    ms = timed(step, steps=10, warmup=4, flush=False)
    return {"config": "C4: MinkUNet-14 shape (8.0 M params), %d scenes, %d voxels/step, AMP bf16, "
                      "fwd+bwd+SGD, kernel maps rebuilt every step" % (scenes, n),
            "ms_per_step": ms, "voxels_per_s": n / (ms * 1e-3),
            "peak_mem_GiB": torch.cuda.max_memory_allocated() / 2 ** 30}


def c5():
    """Group conv (512 ch, groups = 64) on 125k voxels (one rank's share of 1M) and PointConv
    knn_k = 16 on 125k points, bf16, fwd + bwd."""
    c = surface_scene(354, 0).cuda()
    n = len(c)
    bc = torch.cat([torch.zeros(n, 1, dtype=torch.int32, device="cuda"), c], 1).contiguous()
    km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
    x = torch.randn(n, 512, device="cuda").bfloat16()
    gy = torch.randn(n, 512, device="cuda").bfloat16()
    w = (torch.randn(27, 64, 8, 8, device="cuda") * (27 * 8) ** -0.5).bfloat16()

    def gstep():
        sparse_conv_forward(x, w, km, n, groups=64)
        sparse_conv_dgrad(gy, w, km, n, groups=64)
        sparse_conv_wgrad(x, gy, tuple(w.shape), km, groups=64)

    t_group = timed(gstep)
    npts = 125000
    pts = torch.rand(npts, 3, device="cuda")
    feats = torch.randn(npts, 64, device="cuda")
    offs = torch.tensor([0, npts], dtype=torch.int64)
    conv = PointConv(64, 64, RealSearchConfig("knn", knn_k=16)).cuda()
    t_knn = timed(lambda: _ops.knn_search(pts, offs, pts, offs, 16))

    def pstep():
        pc = Points(pts, feats.detach().requires_grad_(True), offsets=offs)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = conv(pc)
        out.feature_tensor.float().square().mean().backward()

    t_pc = timed(pstep, flush=False)
    return {"config": "C5: SparseConv3d(512,512,3,groups=64) on %d voxels fwd+dgrad+wgrad; "
                      "PointConv(64,64,knn_k=16) on %d points fwd+bwd" % (n, npts),
            "group_conv_fwd_bwd_ms": t_group, "group_conv_voxels_per_s": n / (t_group * 1e-3),
            "knn_ms": t_knn, "pointconv_fwd_bwd_ms": t_pc,
            "pointconv_points_per_s": npts / (t_pc * 1e-3)}


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c4", "c5"]
    for name in which:
        t0 = time.time()
        out = globals()[name]()
        out["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(out), flush=True)
