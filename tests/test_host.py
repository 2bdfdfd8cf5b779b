# SPDX-License-Identifier: Apache-2.0
"""CPU tests of the host-side mirror of the reference API (no device compute)."""
import math

import numpy as np
import pytest
import torch

from oracle import kernel_map as okm
from warpconvnet_b200.geometry.coords.ops.stride import stride_coords, unique_coords
from warpconvnet_b200.geometry.coords.search.cache import IntSearchCache, IntSearchCacheKey
from warpconvnet_b200.geometry.coords.search.torch_discrete import kernel_offsets_from_size
from warpconvnet_b200.geometry.types.voxels import Voxels
from warpconvnet_b200.nn.modules.sparse_conv import SparseConv2d, SparseConv3d
from warpconvnet_b200.utils.ntuple import ntuple


def test_ntuple():
    assert ntuple(3, ndim=3) == (3, 3, 3)
    assert ntuple((1, 2, 3), ndim=3) == (1, 2, 3)


def test_kernel_offsets_match_oracle_order():
    for ks in ((3, 3, 3), (2, 2, 2), (5, 3, 1), (1, 1, 1)):
        got = kernel_offsets_from_size(ks, (1, 1, 1)).numpy()
        assert np.array_equal(got[:, 0], np.zeros(len(got), np.int32))  # batch column
        assert np.array_equal(got[:, 1:], okm.kernel_offsets(ks))
    got = kernel_offsets_from_size((3, 3, 3), (2, 1, 3)).numpy()
    assert np.array_equal(got[:, 1:], okm.kernel_offsets((3, 3, 3), (2, 1, 3)))


def test_module_parameters_and_init():
    """weight [K,Cin,Cout] / [K,G,Cin/G,Cout/G], bias [Cout], kaiming-uniform with the
    sqrt(num_spatial_dims) bound (reference nn/modules/sparse_conv.py:147-157,198-217)."""
    torch.manual_seed(0)
    m = SparseConv3d(16, 32, 3)
    assert tuple(m.weight.shape) == (27, 16, 32) and tuple(m.bias.shape) == (32,)
    gain = torch.nn.init.calculate_gain("leaky_relu", math.sqrt(5))
    bound = math.sqrt(3) * gain / math.sqrt(16 * 27)
    assert float(m.weight.abs().max()) <= bound + 1e-7
    assert float(m.weight.abs().max()) > 0.9 * bound
    assert float(m.bias.abs().max()) <= 1 / math.sqrt(16 * 27) + 1e-7
    g = SparseConv3d(64, 128, 3, groups=8, bias=False)
    assert tuple(g.weight.shape) == (27, 8, 8, 16) and g.bias is None
    assert tuple(SparseConv3d(8, 8, 2, stride=2).weight.shape) == (8, 8, 8)
    assert tuple(SparseConv2d(8, 4, 3).weight.shape) == (9, 8, 4)
    with pytest.raises(ValueError):
        SparseConv3d(10, 16, 3, groups=4)
    assert "kernel_size=(3, 3, 3)" in repr(m)


def _toy():
    b0 = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [2, 2, 2], [3, 1, 3], [4, 3, 0]],
                      dtype=torch.int32)
    b1 = torch.tensor([[4, 4, 4], [5, 4, 4], [4, 5, 4], [4, 4, 5], [6, 2, 0], [3, 6, 1], [7, 1, 3]],
                      dtype=torch.int32)
    f0 = torch.arange(21, dtype=torch.float32).reshape(7, 3)
    return Voxels([b0, b1], [f0, f0 + 100])


def test_voxels_container():
    v = _toy()
    assert v.offsets.tolist() == [0, 7, 14] and v.batch_size == 2
    bc = v.batch_indexed_coordinates
    assert bc.shape == (14, 4) and bc.dtype == torch.int32
    assert bc[:, 0].tolist() == [0] * 7 + [1] * 7
    assert v.batch_indexed_coordinates is bc  # cached, not rebuilt per call
    r = v.replace(batched_features=v.feature_tensor * 2)
    assert torch.equal(r.feature_tensor, v.feature_tensor * 2)
    assert r.batched_coordinates is v.batched_coordinates
    assert r._extra_attributes is v._extra_attributes or r.cache is v.cache
    # cat-tensor + offsets constructor
    v2 = Voxels(v.coordinate_tensor, v.feature_tensor, offsets=v.offsets)
    assert torch.equal(v2.batch_indexed_coordinates, bc)


def test_unique_and_stride_coords_match_oracle():
    rng = np.random.RandomState(0)
    c = rng.randint(-9, 9, size=(500, 3)).astype(np.int32)
    bc = okm.batch_indexed([c[:250], c[250:]])
    uniq, idx = unique_coords(torch.from_numpy(bc))
    assert len(np.unique(bc, axis=0)) == len(uniq)
    assert torch.equal(torch.from_numpy(bc)[idx], uniq)
    for stride in ((2, 2, 2), (2, 1, 4)):
        out, offs = stride_coords(torch.from_numpy(bc), stride)
        ref, ref_offs = okm.stride_coords(bc, stride)  # floor division incl. negatives
        assert np.array_equal(out.numpy(), ref)
        assert offs.tolist() == ref_offs.tolist()


def test_search_cache_key():
    a = IntSearchCacheKey((3, 3, 3), (1, 1, 1), False, False, "x", False, torch.tensor([0, 5]),
                          torch.tensor([0, 5]))
    b = IntSearchCacheKey((3, 3, 3), (1, 1, 1), False, False, "x", False, torch.tensor([0, 5]),
                          torch.tensor([0, 5]))
    c = IntSearchCacheKey((3, 3, 3), (1, 1, 1), True, False, "x", False, torch.tensor([0, 5]),
                          torch.tensor([0, 5]))
    assert a == b and hash(a) == hash(b) and a != c
    cache = IntSearchCache()
    cache.put(a, "map")
    assert cache.get(b) == "map" and cache.get(c) is None


def test_pointwise_shortcut_on_cpu_and_no_cpu_fallback_for_kernels():
    """1x1x1 stride-1 is `feats @ weight[0] (+bias)` with no kernel map (helper.py:206-213);
    anything that needs the device kernels must refuse CPU tensors loudly."""
    v = _toy()
    m = SparseConv3d(3, 5, 1)
    out = m(v)
    ref = v.feature_tensor @ m.weight[0] + m.bias
    assert torch.allclose(out.feature_tensor, ref)
    with pytest.raises((RuntimeError, AssertionError)):
        SparseConv3d(3, 5, 3)(v)


def test_batchnorm_module_mirrors_reference_names():
    """BatchNorm keeps the reference's `norm.*` parameter / buffer names
    (warpconvnet/nn/modules/normalizations.py:53-67) so state dicts interchange with
    nn.BatchNorm1d wrapped the reference's way."""
    import torch
    from warpconvnet_b200.nn.modules.normalizations import BatchNorm
    bn = BatchNorm(16, eps=1e-4, momentum=0.05, relu=True)
    keys = set(bn.state_dict().keys())
    assert keys == {"norm.weight", "norm.bias", "norm.running_mean", "norm.running_var",
                    "norm.num_batches_tracked"}
    ref = torch.nn.BatchNorm1d(16, eps=1e-4, momentum=0.05)
    bn.norm.load_state_dict(ref.state_dict())
    assert bn.norm.eps == 1e-4 and bn.norm.momentum == 0.05
    with __import__("pytest").raises(RuntimeError):   # no CPU fallback
        bn(torch.randn(4, 16))


def test_depthwise_module_shapes_and_init_bounds():
    """SparseDepthwiseConv3d: weight [K, C], bias [C], kaiming-uniform bound of the reference
    (sparse_conv_depth.py:143-182): sqrt(3) * gain(leaky_relu, sqrt(5)) / sqrt(K)."""
    import math
    import torch
    from warpconvnet_b200.nn.modules.sparse_conv_depth import SparseDepthwiseConv2d, SparseDepthwiseConv3d
    torch.manual_seed(0)
    m = SparseDepthwiseConv3d(24, 3)
    assert tuple(m.weight.shape) == (27, 24) and tuple(m.bias.shape) == (24,)
    gain = math.sqrt(2.0 / (1 + 5.0))
    bound = math.sqrt(3) * gain / math.sqrt(27)
    assert float(m.weight.abs().max()) <= bound + 1e-7
    assert float(m.bias.abs().max()) <= 1 / math.sqrt(27) + 1e-7
    m2 = SparseDepthwiseConv2d(8, (3, 5), bias=False)
    assert tuple(m2.weight.shape) == (15, 8) and m2.bias is None and m2.num_spatial_dims == 2
    assert "SparseDepthwiseConv3d(channels=24" in repr(m)


def test_radius_config_and_wrappers_reject_cpu():
    import pytest
    import torch
    from warpconvnet_b200.geometry.coords.search.radius import batched_radius_search, radius_search
    from warpconvnet_b200.geometry.coords.search.search_configs import RealSearchConfig, RealSearchMode
    cfg = RealSearchConfig("radius", radius=0.25)
    assert cfg.mode == RealSearchMode.RADIUS and cfg.radius == 0.25
    with pytest.raises(RuntimeError):
        radius_search(torch.rand(10, 3), torch.rand(4, 3), 0.1)
    with pytest.raises(RuntimeError):
        batched_radius_search(torch.rand(10, 3), torch.tensor([0, 10]), torch.rand(4, 3),
                              torch.tensor([0, 4]), 0.1)


def test_to_csr_and_pair_table_reduction_on_cpu():
    """IntSearchResult.to_csr (search_results.py:147-174) and the pair-table pooling walk are plain
    torch: checked on the CPU against a hand-built map."""
    from warpconvnet_b200.geometry.coords.search.search_results import IntSearchResult
    from warpconvnet_b200.nn.functional.sparse_pool import reduce_over_pair_table
    from warpconvnet_b200.ops.reductions import REDUCTIONS
    # offset 0: (in 0 -> out 1), (in 2 -> out 0); offset 1: (in 1 -> out 1), (in 3 -> out 3)
    km = IntSearchResult(torch.tensor([0, 2, 1, 3], dtype=torch.int32),
                         torch.tensor([1, 0, 1, 3], dtype=torch.int32),
                         torch.tensor([0, 2, 4], dtype=torch.int32))
    in_rows, out_rows, offs = km.to_csr()
    assert in_rows.tolist() == [2, 0, 1, 3] and out_rows.tolist() == [0, 1, 3]
    assert offs.tolist() == [0, 1, 3, 4]
    table = torch.tensor([[2, 0, -1, -1], [-1, 1, -1, 3]], dtype=torch.int32)
    x = torch.tensor([[1., -2.], [3., 4.], [-5., 6.], [7., 8.]])
    assert reduce_over_pair_table(x, table, REDUCTIONS.MAX).tolist() == \
        [[-5., 6.], [3., 4.], [0., 0.], [7., 8.]]
    assert reduce_over_pair_table(x, table, REDUCTIONS.MIN)[1].tolist() == [1., -2.]
    assert reduce_over_pair_table(x, table, REDUCTIONS.SUM)[1].tolist() == [4., 2.]
    assert reduce_over_pair_table(x, table, REDUCTIONS.MEAN)[1].tolist() == [2., 1.]
    x.requires_grad_(True)
    reduce_over_pair_table(x, table, REDUCTIONS.MAX).sum().backward()
    assert x.grad.tolist() == [[0., 0.], [1., 1.], [1., 1.], [1., 1.]]


def test_geometry_container_api_matches_reference_surface():
    """Members of Geometry / BatchedTensor / Voxels / Points that the reference exposes
    (geometry/base/geometry.py:37-388, base/batched.py:14-270, types/voxels.py, types/points.py):
    arithmetic, aliases, nested / padded views, dense round trip, z-order sort, downsampling."""
    from warpconvnet_b200.geometry.base.batched import CatFeatures, PadFeatures
    from warpconvnet_b200.geometry.coords.ops.serialization import POINT_ORDERING, morton_code
    from warpconvnet_b200.geometry.types.points import Points
    from warpconvnet_b200.geometry.types.voxels import Voxels
    g = torch.Generator().manual_seed(0)
    coords = [torch.randint(0, 8, (n, 3), generator=g, dtype=torch.int32) for n in (20, 12)]
    feats = [torch.randn(n, 4, generator=g) for n in (20, 12)]
    v = Voxels(coords, feats).unique()
    n = v.coordinate_tensor.shape[0]
    assert len(v) == n and v.coords.shape == (n, 4) and v.coordinates.shape == (n, 3)
    assert torch.equal(v.feats, v.features) and str(v).startswith("Voxels(feature_shape=")
    # arithmetic: geometry, scalar, reflected, per-channel tensor
    assert torch.allclose((v + v).feature_tensor, 2 * v.feature_tensor)
    assert torch.allclose((1 - v).feature_tensor, 1 - v.feature_tensor)
    assert torch.allclose((2 / (v * v + 1)).feature_tensor, 2 / (v.feature_tensor ** 2 + 1))
    assert torch.allclose((v ** 2).feature_tensor, v.feature_tensor ** 2)
    assert torch.allclose((v * torch.arange(4.)).feature_tensor, v.feature_tensor * torch.arange(4.))
    assert v.equal_shape(v * 2)
    with pytest.raises(AssertionError):
        v + Voxels(coords[:1], feats[:1])  # different batch layout
    # batched tensors
    f = v.batched_features
    assert f == f * 2 and f.equal_rigorous(f + 0) and not f.equal_rigorous(f + 1)
    assert len(f) == n and (f - f).batched_tensor.abs().max() == 0
    nested = v.nested_features
    assert nested.is_nested and CatFeatures.from_nested(nested).equal_rigorous(f)
    pad = v.padded_features
    assert isinstance(pad, PadFeatures) and pad.shape[0] == 2 and pad.to_cat().equal_rigorous(f)
    assert v.to_pad(8).batched_features.shape[1] % 8 == 0 and v.to_pad(8).to_cat().to_cat() is not None
    # dense round trip (from_dense lists cells in (b, x, y, z) order = unique()'s order)
    dense = v.to_dense(channel_dim=1, spatial_shape=(8, 8, 8), min_coords=(0, 0, 0))
    back = Voxels.from_dense(dense)
    keep = v.feature_tensor.abs().sum(1) > 0
    assert torch.equal(back.batch_indexed_coordinates, v.batch_indexed_coordinates[keep])
    assert torch.equal(back.feature_tensor, v.feature_tensor[keep])
    assert torch.equal(Voxels.from_dense(dense, target_spatial_sparse_tensor=v).feature_tensor,
                       v.feature_tensor)
    # z-order: code definition (first axis in the lowest bit) and per-scene sortedness
    assert morton_code(torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [3, 0, 0]])
                       ).tolist() == [0, 1, 2, 4, 9]
    s = v.sort()
    assert s.ordering == POINT_ORDERING.MORTON_XYZ and s.offsets.tolist() == v.offsets.tolist()
    codes = morton_code(s.coordinate_tensor)    # normalised by the minimum over the whole batch
    for b in range(2):
        lo, hi = int(s.offsets[b]), int(s.offsets[b + 1])
        code = codes[lo:hi]
        assert bool((code[1:] >= code[:-1]).all())
        assert sorted(map(tuple, s.coordinate_tensor[lo:hi].tolist())) == \
            sorted(map(tuple, v.coordinate_tensor[lo:hi].tolist()))
    # points
    p = v.to_point(voxel_size=0.5)
    assert isinstance(p, Points) and torch.equal(p.coordinate_tensor, v.coordinate_tensor * 0.5)
    pts = Points([torch.rand(50, 3, generator=g), torch.rand(30, 3, generator=g)],
                 [torch.randn(50, 4, generator=g), torch.randn(30, 4, generator=g)])
    down = pts.voxel_downsample(0.25, reduction="mean")
    cells = torch.floor(pts.coordinate_tensor / 0.25)
    first = cells[:50]
    m0 = len({tuple(c) for c in first.tolist()})
    assert down.offsets[1] == m0 and down.voxel_size == 0.25
    c0 = torch.floor(down.coordinate_tensor[0] / 0.25)
    members = (first == c0).all(1)
    assert torch.allclose(down.feature_tensor[0], pts.feature_tensor[:50][members].mean(0), atol=1e-6)
    assert pts.voxel_downsample(0.25).feature_tensor.shape == down.feature_tensor.shape
    rs = pts.random_downsample(10)
    assert rs.offsets.tolist() == [0, 10, 20]
    assert pts.sort(0.25).coordinate_tensor.shape == (80, 3) and pts.contiguous() is pts
    enc = Points.from_list_of_coordinates([torch.rand(5, 3), torch.rand(7, 3)], encoding_channels=4,
                                          encoding_range=1.0)
    assert enc.feature_tensor.shape == (12, 12)


def test_sparse_ops_cat_prune_and_pool_modules():
    """cat / prune of sparse tensors (nn/functional/sparse_ops.py:13-66), IntCoords.prune / sort and
    the pooling module surface (nn/modules/sparse_pool.py:20-92) — host logic only."""
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.functional.sparse_ops import (cat_spatially_sparse_tensors,
                                                           prune_spatially_sparse_tensor)
    from warpconvnet_b200.nn.modules import SparseAvgPool, SparseMaxPool, SparseMinPool, SparsePool
    g = torch.Generator().manual_seed(1)
    coords = [torch.randint(0, 16, (n, 3), generator=g, dtype=torch.int32) for n in (9, 6)]
    a = Voxels(coords, [torch.randn(n, 3, generator=g) for n in (9, 6)])
    b = a.replace(batched_features=torch.randn(15, 5, generator=g))
    c = cat_spatially_sparse_tensors(a, b)
    assert c.feature_tensor.shape == (15, 8)
    assert torch.equal(c.feature_tensor[:, :3], a.feature_tensor)
    assert torch.equal(c.coordinate_tensor, a.coordinate_tensor)
    with pytest.raises(ValueError):
        cat_spatially_sparse_tensors(a, Voxels(coords[:1], [torch.randn(9, 2)]))
    mask = torch.zeros(15, dtype=torch.bool)
    mask[[0, 3, 8, 9, 14]] = True
    p = prune_spatially_sparse_tensor(a, mask)
    assert p.offsets.tolist() == [0, 3, 5] and p.cache is None
    assert torch.equal(p.coordinate_tensor, a.coordinate_tensor[mask])
    assert torch.equal(p.feature_tensor, a.feature_tensor[mask])
    with pytest.raises(ValueError):
        prune_spatially_sparse_tensor(a, mask[:5])
    srt = a.batched_coordinates.sort()
    assert srt.offsets.tolist() == a.offsets.tolist()
    assert sorted(map(tuple, srt.batched_tensor[:9].tolist())) == sorted(map(tuple, coords[0].tolist()))
    assert repr(SparseMaxPool(2, 2)) == "SparseMaxPool(kernel_size=2, stride=2, reduce=max)"
    assert SparseMinPool(2, 2).reduce == "min" and SparseAvgPool(3, 2).reduce == "mean"
    assert SparsePool(2, 2, "sum").stride == 2
