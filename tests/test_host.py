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
