# SPDX-License-Identifier: Apache-2.0
"""Edge cases of the hot path through the public modules (the reference tests the same corners:
tests/coords/test_packed_hashmap.py:188-340 boundary coordinates, tests/nn/test_sparse_conv.py
tiny / batched inputs): single voxels, isolated voxels (centre offset only), an empty scene inside
a batch, coordinates at the limits of the packed key, tiny matrices in the norm / depthwise /
neighbour-search rows."""
import numpy as np
import pytest
import torch

from oracle import conv as oconv
from oracle import kernel_map as okm

pytestmark = pytest.mark.gpu


def _conv_vs_oracle(coords_list, cin, cout, ks=3, stride=1):
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(0)
    feats = [torch.randn(len(c), cin) for c in coords_list]
    x = Voxels([torch.as_tensor(c, dtype=torch.int32).reshape(-1, 3) for c in coords_list], feats,
               device="cuda")
    x.batched_features.batched_tensor.requires_grad_(True)
    conv = SparseConv3d(cin, cout, ks, stride, bias=True).cuda()
    out = conv(x)
    n_in = sum(len(c) for c in coords_list)
    bc = okm.batch_indexed([np.asarray(c, np.int32).reshape(-1, 3) for c in coords_list])
    out_bc = bc if stride == 1 else okm.stride_coords(bc, (stride,) * 3)[0]
    km = okm.generate_kernel_map(bc, out_bc, (stride,) * 3, (ks,) * 3)
    xf = torch.cat(feats) if n_in else torch.zeros(0, cin)
    w = conv.weight.detach().cpu()
    y_ref = oconv.forward(xf, w, km["in_maps"], km["out_maps"], km["offsets"], len(out_bc))
    y_ref = y_ref + conv.bias.detach().cpu().double()
    assert out.feature_tensor.shape == (len(out_bc), cout)
    if len(out_bc):
        assert np.array_equal(out.batch_indexed_coordinates.cpu().numpy(), out_bc)
        assert oconv.rel_max_err(out.feature_tensor, y_ref) < 5e-3      # fp32 in, TF32 tensor cores
        out.feature_tensor.sum().backward()
        gy = torch.ones(len(out_bc), cout)
        dx_ref, dw_ref = oconv.backward(gy, xf, w, km["in_maps"], km["out_maps"], km["offsets"])
        assert oconv.rel_max_err(x.batched_features.batched_tensor.grad, dx_ref) < 5e-3
        assert oconv.rel_max_err(conv.weight.grad, dw_ref) < 1e-3
    return out


def test_single_voxel():
    _conv_vs_oracle([[[5, 6, 7]]], 16, 32)


def test_isolated_voxels_only_centre_offset():
    c = [[i * 10, (i * 7) % 50, (i * 3) % 40] for i in range(300)]
    out = _conv_vs_oracle([c], 32, 32)
    assert out.feature_tensor.shape[0] == 300


def test_batch_with_tiny_scenes():
    rng = np.random.RandomState(0)
    big = np.unique(rng.randint(0, 20, size=(3000, 3)), axis=0)
    _conv_vs_oracle([big, [[1, 1, 1]], [[0, 0, 0], [0, 0, 1]]], 16, 48)


def test_coordinates_at_key_limits():
    lim = 131071
    c = [[lim, lim, lim], [lim - 1, lim, lim], [-131072, -131072, -131072], [-131071, -131072, -131072],
         [0, 0, 0], [lim, -131072, 0]]
    _conv_vs_oracle([c], 16, 16)


def test_strided_conv_tiny():
    c = [[0, 0, 0], [1, 1, 1], [2, 2, 2], [3, 3, 3], [9, 9, 9]]
    _conv_vs_oracle([c], 16, 32, ks=2, stride=2)


def test_out_of_range_coordinate_raises():
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    x = Voxels([torch.tensor([[0, 0, 0], [131072, 0, 0]], dtype=torch.int32)], [torch.randn(2, 16)],
               device="cuda")
    conv = SparseConv3d(16, 16, 3).cuda()
    with pytest.raises(ValueError):
        out = conv(x)
        _ = out.feature_tensor.cpu()
        # the status word is read when the kernel map's host side is first needed
        next(iter(x.cache.values())).offsets  # noqa: B018


def test_norm_tiny_and_wide():
    from warpconvnet_b200.nn.functional.normalizations import batch_norm_act
    torch.manual_seed(11)  # 3 samples per channel: an unlucky draw has channels with var ~ 1e-4 mean^2
    for n, c in ((2, 8), (3, 1024), (5, 2)):
        x = torch.randn(n, c).cuda().requires_grad_(True)
        y = batch_norm_act(x, None, None, None, None, training=True, relu=False)
        ref = torch.nn.functional.batch_norm(x.detach().double().cpu(), None, None, None, None, True)
        assert oconv.rel_max_err(y, ref) < 1e-4
        y.sum().backward()
        assert torch.isfinite(x.grad).all()


def test_depthwise_kernel_size_one_and_single_row():
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules import SparseDepthwiseConv3d
    x = Voxels([torch.tensor([[1, 2, 3]], dtype=torch.int32)], [torch.randn(1, 8)], device="cuda")
    conv = SparseDepthwiseConv3d(8, 3, bias=False).cuda()
    out = conv(x)
    ref = x.feature_tensor * conv.weight[13].unsqueeze(0)     # only the centre offset hits
    assert torch.allclose(out.feature_tensor, ref, atol=1e-6)


def test_radius_search_no_neighbours_and_all_neighbours():
    from warpconvnet_b200.geometry.coords.search.radius import batched_radius_search
    ref = torch.tensor([[0., 0, 0], [1, 0, 0], [0, 1, 0]]).cuda()
    q = torch.tensor([[5., 5, 5], [0.1, 0.1, 0]]).cuda()
    off_r, off_q = torch.tensor([0, 3]), torch.tensor([0, 2])
    idx, dist, splits = batched_radius_search(ref, off_r, q, off_q, 0.05)
    assert idx.numel() == 0 and splits.tolist() == [0, 0, 0]
    idx, dist, splits = batched_radius_search(ref, off_r, q, off_q, 100.0)
    assert splits.tolist() == [0, 3, 6] and sorted(idx[:3].tolist()) == [0, 1, 2]
