# SPDX-License-Identifier: Apache-2.0
"""BASELINE config 4 shape: a MinkUNet-14-style network (tools/minkunet14.py) runs forward +
backward on the device path; one encoder level is cross-checked against the CPU oracle."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
pytestmark = pytest.mark.gpu


def test_minkunet14_forward_backward_small():
    from minkunet14 import MinkUNet14, surface_scene
    from warpconvnet_b200.geometry.types.voxels import Voxels
    torch.manual_seed(0)
    coords = [surface_scene(96, s) for s in (0, 1)]
    feats = [torch.randn(len(c), 3) for c in coords]
    x = Voxels(coords, feats, device="cuda")
    net = MinkUNet14(3, 20).cuda()
    import warpconvnet_b200.nn.functional.sparse_conv.helper as helper
    builds = []
    orig = helper.generate_kernel_map
    helper.generate_kernel_map = lambda *a, **k: (builds.append(1), orig(*a, **k))[1]
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(x)
    finally:
        helper.generate_kernel_map = orig
    n = sum(len(c) for c in coords)
    assert out.feature_tensor.shape == (n, 20)
    assert out.offsets.tolist() == x.offsets.tolist()
    loss = out.feature_tensor.float().square().mean()
    loss.backward()
    torch.cuda.synchronize()
    grads = [p.grad for p in net.parameters()]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    assert sum(float(g.abs().sum()) for g in grads) > 0
    # 24 sparse convs, but only 9 kernel maps: one 3^3 map per resolution (5) + one 2^3 stride-2
    # map per level change (4); the transposed convs reuse the encoder maps with in/out swapped
    # (reference helper.py:446-497 caches the same way)
    assert len(builds) == 9


def test_encoder_level_matches_oracle():
    """conv(2^3, stride 2) -> conv(3^3) chain vs the CPU oracle on the same bf16-rounded weights."""
    from minkunet14 import surface_scene
    from oracle import conv as oconv
    from oracle import kernel_map as okm
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(1)
    c = surface_scene(128, 3)
    f = torch.randn(len(c), 32)
    x = Voxels([c], [f], device="cuda")
    down = SparseConv3d(32, 32, 2, 2, bias=False).cuda()
    conv = SparseConv3d(32, 64, 3, bias=False).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        d = down(x)
        y = conv(d)
    bc = okm.batch_indexed([c.numpy()])
    out_bc, _ = okm.stride_coords(bc, (2, 2, 2))
    km1 = okm.generate_kernel_map(bc, out_bc, (2, 2, 2), (2, 2, 2))
    km2 = okm.generate_kernel_map(out_bc, out_bc, (1, 1, 1), (3, 3, 3))
    w1 = down.weight.detach().cpu().bfloat16().float()
    w2 = conv.weight.detach().cpu().bfloat16().float()
    r1 = oconv.forward(f.bfloat16().float(), w1, km1["in_maps"], km1["out_maps"], km1["offsets"],
                       len(out_bc))
    assert np.array_equal(d.batch_indexed_coordinates.cpu().numpy(), out_bc)
    assert oconv.rel_max_err(d.feature_tensor, r1) < 1e-2
    r2 = oconv.forward(d.feature_tensor.float().cpu(), w2, km2["in_maps"], km2["out_maps"],
                       km2["offsets"], len(out_bc))
    assert oconv.rel_max_err(y.feature_tensor, r2) < 1e-2
