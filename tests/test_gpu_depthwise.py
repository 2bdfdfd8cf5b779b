# SPDX-License-Identifier: Apache-2.0
"""GPU parity of the depthwise sparse conv (csrc/conv_depthwise.cu through the C-ABI) against the
CPU oracle and the reference's golden outputs."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import random_coords, surface_coords
from oracle import conv as oconv

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dw_*.npz")))
TOL = {torch.bfloat16: 1e-2, torch.float16: 2e-3, torch.float32: 2e-5}


def _bc(c):
    return torch.from_numpy(np.concatenate([np.zeros((len(c), 1), np.int32), c], 1)).cuda()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_depthwise_vs_reference_golden(path):
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    from warpconvnet_b200.nn.functional.sparse_conv_depth import spatially_sparse_depthwise_conv
    d = np.load(path)
    in_bc, out_bc = torch.from_numpy(d["in_bc"]).cuda(), torch.from_numpy(d["out_bc"]).cuda()
    stride, ks = tuple(int(s) for s in d["stride"]), tuple(int(k) for k in d["ksize"])
    same = stride == (1, 1, 1)
    km = generate_kernel_map(in_bc, in_bc if same else out_bc, stride, ks, same_coords=same)
    assert np.array_equal(km.offsets.numpy(), d["offsets"])
    x = torch.from_numpy(d["x"]).float().cuda().requires_grad_(True)
    w = torch.from_numpy(d["w"]).float().cuda().requires_grad_(True)
    y = spatially_sparse_depthwise_conv(x, w, km, len(d["out_bc"]))
    y.backward(torch.from_numpy(d["gy"]).float().cuda())
    assert oconv.rel_max_err(y, torch.from_numpy(d["y"])) < 2e-5
    assert oconv.rel_max_err(x.grad, torch.from_numpy(d["dx"])) < 2e-5
    assert oconv.rel_max_err(w.grad, torch.from_numpy(d["dw"])) < 2e-5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("kind,n,c,stride,ks", [("R", 20000, 32, 1, 3), ("S", 120, 128, 1, 3),
                                                 ("R", 30000, 20, 2, 2), ("R", 8000, 96, 2, 3),
                                                 ("R", 3000, 7, 1, 5)])
def test_depthwise_vs_oracle(dtype, kind, n, c, stride, ks):
    from oracle import kernel_map as okm
    from warpconvnet_b200.geometry.coords.ops.stride import stride_coords
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    from warpconvnet_b200.nn.functional.sparse_conv_depth import spatially_sparse_depthwise_conv
    coords = surface_coords(n, 2) if kind == "S" else random_coords(n, 0.3, 2)
    bc = _bc(coords)
    out_bc = bc if stride == 1 else stride_coords(bc, (stride,) * 3)[0]
    km = generate_kernel_map(bc, out_bc, (stride,) * 3, (ks,) * 3, same_coords=stride == 1)
    K, m = ks ** 3, len(out_bc)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(len(bc), c, generator=g).to(dtype).cuda().requires_grad_(True)
    w = (torch.randn(K, c, generator=g) * K ** -0.5).cuda().requires_grad_(True)
    gy = torch.randn(m, c, generator=g).to(dtype).cuda()
    y = spatially_sparse_depthwise_conv(x, w, km, m)
    assert y.dtype == dtype and y.shape == (m, c)
    y.backward(gy)
    args = (km.in_maps.cpu().numpy(), km.out_maps.cpu().numpy(), km.offsets.numpy())
    y_ref = oconv.depthwise_forward(x.detach().float().cpu(), w.detach().cpu(), *args, m)
    dx_ref, dw_ref = oconv.depthwise_backward(gy.float().cpu(), x.detach().float().cpu(),
                                              w.detach().cpu(), *args)
    tol = TOL[dtype]
    assert oconv.rel_max_err(y, y_ref) < tol
    assert oconv.rel_max_err(x.grad, dx_ref) < tol
    assert oconv.rel_max_err(w.grad, dw_ref) < 1e-4      # fp32 accumulation of exact products


def test_depthwise_module_autocast_and_bias():
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules import SparseDepthwiseConv3d
    torch.manual_seed(0)
    coords = torch.from_numpy(random_coords(5000, 0.3, 3))
    feats = torch.randn(len(coords), 64)
    x = Voxels([coords], [feats], device="cuda")
    conv = SparseDepthwiseConv3d(64, 3, bias=True).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = conv(x)
    assert out.feature_tensor.shape == (len(coords), 64)
    out.feature_tensor.float().sum().backward()
    assert conv.weight.grad is not None and conv.weight.grad.shape == (27, 64)
    assert conv.bias.grad is not None and torch.isfinite(conv.weight.grad).all()
    down = SparseDepthwiseConv3d(64, 2, stride=2, bias=False).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        d = down(x)
    assert d.feature_tensor.shape[1] == 64 and d.feature_tensor.shape[0] < len(coords)
    assert d.tensor_stride == (2, 2, 2)


def test_depthwise_rejects_cpu():
    from warpconvnet_b200.nn.functional.sparse_conv_depth import (
        UnifiedSpatiallySparseDepthwiseConvFunction as F)
    with pytest.raises(RuntimeError):
        F.apply(torch.randn(4, 8), torch.randn(27, 8), None, 4, None)


@pytest.mark.parametrize("c", [32, 96, 20])
def test_depthwise_table_and_plan_paths_agree(c):
    """The dense-table kernel (wcn_depthwise_conv) and the tile-plan kernel
    (wcn_depthwise_conv_plan) compute the same rows; with kflip both equal the reverse-table dgrad."""
    from warpconvnet_b200 import _ops
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    bc = _bc(random_coords(25000, 0.3, 4))
    n = len(bc)
    km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
    g = torch.Generator().manual_seed(c)
    x = torch.randn(n, c, generator=g).bfloat16().cuda()
    w = (torch.randn(27, c, generator=g) * 0.2).cuda()
    b = torch.randn(c, generator=g).cuda()
    table, plan = km.pair_table(n), km.fwd_plan(n)
    for kflip in (False, True):
        a = _ops.depthwise_conv(x, w, table, bias=b, kflip=kflip, relu=True)
        p = _ops.depthwise_conv_plan(x, w, plan, bias=b, kflip=kflip, relu=True)
        assert float((a.float() - p.float()).abs().max()) <= 2e-2 * float(a.float().abs().max())
    gy = torch.randn(n, c, generator=g).bfloat16().cuda()
    dw_t = _ops.depthwise_wgrad(x, gy, table)
    dw_p = _ops.depthwise_wgrad_plan(x, gy, plan)
    assert float((dw_t - dw_p).abs().max()) <= 1e-4 * float(dw_t.abs().max())
    rev = _ops.depthwise_conv(x, w, km.rev_pair_table(n))
    flip = _ops.depthwise_conv_plan(x, w, plan, kflip=True)
    assert float((rev.float() - flip.float()).abs().max()) <= 2e-2 * float(rev.float().abs().max())
