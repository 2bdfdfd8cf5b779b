# SPDX-License-Identifier: Apache-2.0
"""GPU parity tests: the sm_100a path (through the C-ABI in libwcn_b200.so) against the CPU oracle,
the committed golden fixtures, and size-independent properties at BASELINE sizes.

Bars: kernel map bit-exact (offsets, pair_table, CSR in ascending-row order);
features vs the fp64 oracle on the SAME bf16/fp16-rounded operands:
  bf16 / fp16 output: max|d|/max|ref| < 1e-2 (one output rounding, fp32 accumulation) — the
  reference's own bars are 2e-2 (fp16) and mean-relative 1e-1 (bf16);
  fp32 (TF32 tensor cores): < 5e-3;  wgrad (fp32 output): < 1e-4 on bf16 operands.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import random_coords, surface_coords
from oracle import conv as oconv
from oracle import kernel_map as okm

pytestmark = pytest.mark.gpu

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if os.path.basename(p).startswith(("c1_", "toy_")))  # dw_* / radius_*: other rows
TOL = {torch.bfloat16: 1e-2, torch.float16: 1e-2, torch.float32: 5e-3}


def _gkm(in_bc, out_bc, stride, ksize, same=None):
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    it = torch.from_numpy(np.ascontiguousarray(in_bc)).cuda()
    ot = it if out_bc is in_bc else torch.from_numpy(np.ascontiguousarray(out_bc)).cuda()
    return generate_kernel_map(it, ot, tuple(stride), tuple(ksize))


def _assert_map_equal(km, ref):
    assert np.array_equal(km.offsets.numpy(), ref["offsets"])
    assert np.array_equal(km._pair_table.cpu().numpy(), ref["pair_table"])
    assert np.array_equal(km.in_maps.cpu().numpy(), ref["in_maps"])
    assert np.array_equal(km.out_maps.cpu().numpy(), ref["out_maps"])


# ------------------------------------------------------------------------------------------------
# hash table
# ------------------------------------------------------------------------------------------------
def test_hash_roundtrip_dedup_and_range():
    from warpconvnet_b200.geometry.coords.search.packed_hashmap import PackedHashTable
    bc = okm.batch_indexed([random_coords(5000, 0.3, 0), random_coords(3000, 0.3, 1)])
    t = torch.from_numpy(bc).cuda()
    tab = PackedHashTable.from_coords(t)
    assert tab.capacity == 16384
    assert np.array_equal(tab.search(t).cpu().numpy(), np.arange(len(bc)))
    miss = t.clone(); miss[:, 1] += 5000
    assert (tab.search(miss) == -1).all()
    # insertion index is the value (tests/coords/test_packed_hashmap.py:105-114)
    small = torch.tensor([[0, 1, 2, 3], [0, 4, 5, 6], [1, 1, 2, 3]], dtype=torch.int32).cuda()
    assert PackedHashTable.from_coords(small).search(small).tolist() == [0, 1, 2]
    # duplicates keep the smallest index
    dup = torch.tensor([[0, 1, 2, 3], [0, 1, 2, 3], [0, 9, 9, 9], [0, 1, 2, 3]], dtype=torch.int32).cuda()
    assert PackedHashTable.from_coords(dup).search(dup).tolist() == [0, 0, 2, 0]
    # boundary coordinates
    edge = torch.tensor([[511, -131072, 131071, -1], [0, 131071, -131072, 0]], dtype=torch.int32).cuda()
    assert PackedHashTable.from_coords(edge).search(edge).tolist() == [0, 1]
    with pytest.raises(ValueError):
        PackedHashTable.from_coords(torch.tensor([[512, 0, 0, 0]], dtype=torch.int32).cuda())
    with pytest.raises(ValueError):
        PackedHashTable.from_coords(torch.tensor([[0, 0, 131072, 0]], dtype=torch.int32).cuda())


# ------------------------------------------------------------------------------------------------
# kernel map
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_kernel_map_golden(path):
    g = np.load(path)
    same = np.array_equal(g["in_bcoords"], g["out_bcoords"])
    in_bc = g["in_bcoords"]
    km = _gkm(in_bc, in_bc if same else g["out_bcoords"], g["stride"], g["kernel_size"])
    _assert_map_equal(km, g)
    iden = -1 if km.identity_map_index is None else km.identity_map_index
    assert iden == int(g["identity_map_index"])


@pytest.mark.parametrize("n,stride,ks", [(20000, 1, 3), (50000, 2, 2), (30000, 2, 3), (3000, 1, 5),
                                         (257, 1, 3), (1, 1, 3)])
def test_kernel_map_vs_oracle(n, stride, ks):
    bc = okm.batch_indexed([random_coords(n, 0.3, 0), random_coords(max(n // 2, 1), 0.2, 1)])
    out_bc = bc if stride == 1 else okm.stride_coords(bc, (stride,) * 3)[0]
    km = _gkm(bc, out_bc, (stride,) * 3, (ks,) * 3)
    _assert_map_equal(km, okm.generate_kernel_map(bc, out_bc, (stride,) * 3, (ks,) * 3))


def test_kernel_map_negative_coords_and_dilation():
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    c = random_coords(4000, 0.3, 3) - 7
    bc = okm.batch_indexed([c])
    t = torch.from_numpy(bc).cuda()
    km = generate_kernel_map(t, t, (1, 1, 1), (3, 3, 3), kernel_dilation=(2, 1, 2))
    ref = okm.generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), dilation=(2, 1, 2))
    _assert_map_equal(km, ref)


def test_kernel_map_full_size_invariants():
    """BASELINE C3 size (200k voxels, both distributions): in = out + offset[k] for every pair,
    per-offset counts symmetric (L_k == L_{K-1-k}), identity offset is the full diagonal, degree
    histogram equals the pair-table popcount, and offsets match the oracle."""
    for name, c in (("S", surface_coords(448, 0)), ("R", random_coords(200000, 0.3, 0))):
        bc = okm.batch_indexed([c])
        km = _gkm(bc, bc, (1, 1, 1), (3, 3, 3))
        offs = okm.kernel_offsets((3, 3, 3))
        im, om, off = km.in_maps.cpu().numpy(), km.out_maps.cpu().numpy(), km.offsets.numpy()
        counts = np.diff(off)
        assert np.array_equal(counts, counts[::-1]), name
        assert counts[13] == len(bc) and np.array_equal(im[off[13]:off[14]], np.arange(len(bc)))
        kidx = np.repeat(np.arange(27), counts)
        assert np.array_equal(bc[im][:, 1:], bc[om][:, 1:] + offs[kidx]), name
        for k in range(27):
            assert np.all(np.diff(om[off[k]:off[k + 1]]) > 0)  # ascending rows: deterministic CSR
        pt = km._pair_table.cpu().numpy()
        assert np.array_equal((pt >= 0).sum(1), counts)
        ref = okm.generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3))
        _assert_map_equal(km, ref)


def test_kernel_map_empty():
    e = np.zeros((0, 4), np.int32)
    km = _gkm(e, e, (1, 1, 1), (3, 3, 3))
    assert km.offsets.tolist() == [0] * 28 and km.in_maps.numel() == 0


# ------------------------------------------------------------------------------------------------
# the three GEMMs
# ------------------------------------------------------------------------------------------------
def _case(n, cin, cout, dtype, stride=1, ks=3, seed=0, dist="R"):
    c = random_coords(n, 0.3, seed) if dist == "R" else surface_coords(int(n ** 0.5), seed)
    bc = okm.batch_indexed([c])
    out_bc = bc if stride == 1 else okm.stride_coords(bc, (stride,) * 3)[0]
    km = _gkm(bc, out_bc, (stride,) * 3, (ks,) * 3)
    K = ks ** 3
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(len(bc), cin, generator=g).cuda().to(dtype)
    w = (torch.randn(K, cin, cout, generator=g) * (K * cin) ** -0.5).cuda().to(dtype)
    gy = torch.randn(len(out_bc), cout, generator=g).cuda().to(dtype)
    return km, x, w, gy, len(bc), len(out_bc)


def _oracle(km, x, w, gy, n_out):
    args = (km.in_maps.cpu().numpy(), km.out_maps.cpu().numpy(), km.offsets.numpy())
    y = oconv.forward(x.float().cpu(), w.float().cpu(), *args, n_out)
    dx, dw = oconv.backward(gy.float().cpu(), x.float().cpu(), w.float().cpu(), *args)
    return y, dx, dw


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("n,cin,cout,stride,ks", [
    (4000, 64, 128, 1, 3), (20000, 128, 128, 1, 3), (20000, 32, 64, 2, 2), (6000, 96, 32, 2, 3),
    (3000, 16, 16, 1, 3), (3000, 256, 256, 1, 3), (2000, 64, 512, 1, 3), (130, 32, 32, 1, 3),
])
def test_three_gemms_vs_oracle(dtype, n, cin, cout, stride, ks):
    from warpconvnet_b200.nn.functional.sparse_conv import (sparse_conv_dgrad,
                                                            sparse_conv_forward, sparse_conv_wgrad)
    km, x, w, gy, n_in, n_out = _case(n, cin, cout, dtype, stride, ks)
    y = sparse_conv_forward(x, w, km, n_out)
    dx = sparse_conv_dgrad(gy, w, km, n_in)
    dw = sparse_conv_wgrad(x, gy, tuple(w.shape), km)
    torch.cuda.synchronize()
    assert y.dtype == dtype and dx.dtype == dtype and dw.dtype == torch.float32
    ry, rdx, rdw = _oracle(km, x, w, gy, n_out)
    assert oconv.rel_max_err(y, ry) < TOL[dtype]
    assert oconv.rel_max_err(dx, rdx) < TOL[dtype]
    assert oconv.rel_max_err(dw, rdw) < (5e-3 if dtype == torch.float32 else 1e-4)


@pytest.mark.parametrize("cin,cout", [(4, 8), (3, 20), (48, 96), (192, 64)])
def test_odd_channel_counts_are_padded(cin, cout):
    from warpconvnet_b200.nn.functional.sparse_conv import (sparse_conv_dgrad,
                                                            sparse_conv_forward, sparse_conv_wgrad)
    km, x, w, gy, n_in, n_out = _case(3000, cin, cout, torch.bfloat16)
    ry, rdx, rdw = _oracle(km, x, w, gy, n_out)
    assert oconv.rel_max_err(sparse_conv_forward(x, w, km, n_out), ry) < 1e-2
    assert oconv.rel_max_err(sparse_conv_dgrad(gy, w, km, n_in), rdx) < 1e-2
    assert oconv.rel_max_err(sparse_conv_wgrad(x, gy, tuple(w.shape), km), rdw) < 1e-4


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_gemms_vs_reference_golden_outputs(path):
    """fp32 inputs (TF32 tensor cores) against the outputs of the reference's own explicit path."""
    from warpconvnet_b200.nn.functional.sparse_conv import (sparse_conv_dgrad,
                                                            sparse_conv_forward, sparse_conv_wgrad)
    g = np.load(path)
    same = np.array_equal(g["in_bcoords"], g["out_bcoords"])
    km = _gkm(g["in_bcoords"], g["in_bcoords"] if same else g["out_bcoords"], g["stride"],
              g["kernel_size"])
    x, w, gy = (torch.from_numpy(g[k]).cuda() for k in ("x", "w", "gy"))
    n_in, n_out = len(g["in_bcoords"]), len(g["out_bcoords"])
    assert oconv.rel_max_err(sparse_conv_forward(x, w, km, n_out), g["y_ref_f64"]) < 5e-3
    assert oconv.rel_max_err(sparse_conv_dgrad(gy, w, km, n_in), g["dx_ref_f64"]) < 5e-3
    assert oconv.rel_max_err(sparse_conv_wgrad(x, gy, tuple(w.shape), km), g["dw_ref_f64"]) < 5e-3


def test_ones_kat_full_size():
    """x = 1, w = 1 => Y[r, :] = Cin * degree(r) EXACTLY (small integers are exact in bf16 and
    fp32 accumulation) at N = 200k, repeated 3x to catch intermittent races
    (reference scripts/validate_tiles_on_device.py:46-96,165-176)."""
    from warpconvnet_b200.nn.functional.sparse_conv import (sparse_conv_dgrad,
                                                            sparse_conv_forward, sparse_conv_wgrad)
    c = surface_coords(448, 0)
    bc = okm.batch_indexed([c])
    km = _gkm(bc, bc, (1, 1, 1), (3, 3, 3))
    n, cin, cout = len(bc), 128, 128
    x = torch.ones(n, cin, device="cuda", dtype=torch.bfloat16)
    w = torch.ones(27, cin, cout, device="cuda", dtype=torch.bfloat16)
    deg = torch.from_numpy(oconv.degree(km.out_maps.cpu().numpy(), n)).cuda()
    rowv = (torch.arange(n, device="cuda") % 7 + 1).to(torch.bfloat16)  # row-varying gradient
    gy = rowv[:, None].expand(n, cout).contiguous()
    counts = np.diff(km.offsets.numpy())
    for _ in range(3):
        y = sparse_conv_forward(x, w, km, n)
        assert torch.equal(y.float(), (cin * deg).float()[:, None].expand(n, cout))
        dx = sparse_conv_dgrad(gy, w, km, n)
        exp = torch.zeros(n, device="cuda")
        exp.index_add_(0, km.in_maps.long(), rowv.float()[km.out_maps.long()])
        assert torch.equal(dx.float(), (cout * exp).to(torch.bfloat16).float()[:, None].expand(n, cin))
        dw = sparse_conv_wgrad(x, gy, (27, cin, cout), km)
        sums = torch.zeros(27, device="cuda", dtype=torch.float64)
        kidx = torch.from_numpy(np.repeat(np.arange(27), counts)).cuda()
        sums.index_add_(0, kidx, rowv.double()[km.out_maps.long()])
        assert torch.equal(dw.double(), sums[:, None, None].expand(27, cin, cout))


def test_full_size_properties_c3():
    """BASELINE C3 (128->128, 200k voxels, bf16): linearity in x, agreement of a row sample with
    the fp64 oracle, and <dY, conv(X)> == <dgrad(dY), X> == <W, wgrad(X, dY)> (adjoint identity
    ties the three GEMMs together)."""
    from warpconvnet_b200.nn.functional.sparse_conv import (sparse_conv_dgrad,
                                                            sparse_conv_forward, sparse_conv_wgrad)
    km, x, w, gy, n_in, n_out = _case(200704, 128, 128, torch.bfloat16, dist="S")
    y = sparse_conv_forward(x, w, km, n_out)
    y2 = sparse_conv_forward(x * 2, w, km, n_out)
    assert torch.equal(y2.float(), (y.float() * 2))  # exact: power-of-two scaling
    dx = sparse_conv_dgrad(gy, w, km, n_in)
    dw = sparse_conv_wgrad(x, gy, tuple(w.shape), km)
    # rows sample vs oracle
    pt = km._pair_table.cpu().numpy()
    rows = np.random.RandomState(0).choice(n_out, 512, replace=False)
    xf, wf = x.double().cpu(), w.double().cpu()
    ref = torch.zeros(512, 128, dtype=torch.float64)
    for k in range(27):
        src = pt[k, rows]
        ok = src >= 0
        ref[ok] += xf[src[ok]] @ wf[k]
    assert oconv.rel_max_err(y[rows], ref) < 1e-2
    # adjoint identities with fp64 accumulation of the device outputs
    a = float((gy.double() * y.double()).sum())
    b = float((dx.double() * x.double()).sum())
    cc = float((dw.double() * w.double()).sum())
    scale = float(gy.double().abs().mul(y.double().abs()).sum())
    assert abs(a - b) / scale < 2e-3 and abs(a - cc) / scale < 2e-3


@pytest.mark.parametrize("dist,n,cin,cout", [("S", 200704, 128, 128), ("R", 200000, 128, 128),
                                             ("S", 100489, 64, 128)])
def test_full_size_all_rows_vs_fp64_oracle(dist, n, cin, cout):
    """BASELINE C3 (S and R) and C2 at FULL size, EVERY row: forward, dgrad and wgrad against the
    fp64 CPU oracle (oracle/conv.py, the reference's explicit gather-matmul-scatter) on the same
    bf16-rounded operands. The kernel map handed to the oracle is the device's own CSR, which
    test_kernel_map_full_size_invariants pins bit-exact against the NumPy kernel-map oracle."""
    from warpconvnet_b200.nn.functional.sparse_conv import (sparse_conv_dgrad,
                                                            sparse_conv_forward, sparse_conv_wgrad)
    km, x, w, gy, n_in, n_out = _case(n, cin, cout, torch.bfloat16, dist=dist)
    y = sparse_conv_forward(x, w, km, n_out)
    dx = sparse_conv_dgrad(gy, w, km, n_in)
    dw = sparse_conv_wgrad(x, gy, tuple(w.shape), km)
    im, om, offs = km.in_maps.cpu().numpy(), km.out_maps.cpu().numpy(), km.offsets.numpy()
    xf, wf, gf = x.float().cpu(), w.float().cpu(), gy.float().cpu()
    y_ref = oconv.forward(xf, wf, im, om, offs, n_out)
    dx_ref, dw_ref = oconv.backward(gf, xf, wf, im, om, offs)
    assert y.shape == y_ref.shape and dx.shape == dx_ref.shape
    assert oconv.rel_max_err(y, y_ref) < 1e-2          # bf16 output rounding: 2^-9 of max|ref|
    assert oconv.rel_max_err(dx, dx_ref) < 1e-2
    assert oconv.rel_max_err(dw, dw_ref) < 1e-4         # fp32 accumulation of exact bf16 products
    # mean relative error (the reference's own bar is 1e-1 for bf16, tests/nn/test_mask_gemm_numerical.py)
    for got, ref in ((y, y_ref), (dx, dx_ref)):
        d = (got.double().cpu() - ref).abs().mean() / ref.abs().mean()
        assert float(d) < 5e-3


# ------------------------------------------------------------------------------------------------
# module / autograd / groups / transposed / cache
# ------------------------------------------------------------------------------------------------
def _voxels(n_per_scene=(3000, 2500), cin=32, seed=0):
    from warpconvnet_b200.geometry.types.voxels import Voxels
    g = torch.Generator().manual_seed(seed)
    coords = [torch.from_numpy(random_coords(n, 0.3, seed + i)) for i, n in enumerate(n_per_scene)]
    feats = [torch.randn(n, cin, generator=g) for n in n_per_scene]
    return Voxels(coords, feats, device="cuda"), coords, feats


def test_module_forward_backward_autocast():
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(0)
    v, coords, feats = _voxels()
    v.batched_features.batched_tensor.requires_grad_(True)
    conv = SparseConv3d(32, 64, 3).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = conv(v)
    assert out.feature_tensor.dtype == torch.bfloat16 and out.feature_tensor.shape == (5500, 64)
    gy = torch.randn(5500, 64, device="cuda")
    out.feature_tensor.float().backward(gy)
    bc = okm.batch_indexed([c.numpy() for c in coords])
    km = okm.generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3))
    x = torch.cat(feats).bfloat16().float()
    w = conv.weight.detach().cpu().bfloat16().float()
    args = (km["in_maps"], km["out_maps"], km["offsets"])
    ry = oconv.forward(x, w, *args, 5500) + conv.bias.detach().cpu().double()
    rdx, rdw = oconv.backward(gy.cpu().bfloat16().float(), x, w, *args)
    assert oconv.rel_max_err(out.feature_tensor, ry) < 1e-2
    assert oconv.rel_max_err(v.batched_features.batched_tensor.grad, rdx) < 1e-2
    assert oconv.rel_max_err(conv.weight.grad, rdw) < 1e-3
    assert oconv.rel_max_err(conv.bias.grad, gy.double().sum(0).cpu()) < 1e-2
    # same-resolution convs share one kernel map through the Voxels cache
    assert len(v.cache) == 1
    conv2 = SparseConv3d(64, 32, 3).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        conv2(out)
    assert len(out.cache) == 1


def test_strided_and_transposed_modules():
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(1)
    v, coords, feats = _voxels(cin=32)
    down = SparseConv3d(32, 64, 2, stride=2, bias=False).cuda()
    up = SparseConv3d(64, 32, 2, stride=2, transposed=True, bias=False).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        d = down(v)
        u = up(d, v)
    assert d.tensor_stride == (2, 2, 2) and u.feature_tensor.shape == (5500, 32)
    bc = okm.batch_indexed([c.numpy() for c in coords])
    out_bc, offs = okm.stride_coords(bc, (2, 2, 2))
    assert np.array_equal(d.batch_indexed_coordinates.cpu().numpy(), out_bc)
    assert d.offsets.tolist() == offs.tolist()
    km = okm.generate_kernel_map(bc, out_bc, (2, 2, 2), (2, 2, 2))
    x = torch.cat(feats).bfloat16().float()
    wd = down.weight.detach().cpu().bfloat16().float()
    rd = oconv.forward(x, wd, km["in_maps"], km["out_maps"], km["offsets"], len(out_bc))
    assert oconv.rel_max_err(d.feature_tensor, rd) < 1e-2
    # transposed = forward map with in/out swapped (reference helper.py:462-497)
    wu = up.weight.detach().cpu().bfloat16().float()
    ru = oconv.forward(d.feature_tensor.float().cpu(), wu, km["out_maps"], km["in_maps"],
                       km["offsets"], len(bc))
    assert oconv.rel_max_err(u.feature_tensor, ru) < 1e-2


def _pool_oracle(bc, x, stride, how):
    """Dictionary pooling over stride-sized windows (reference sparse_pool.py:25-117 semantics)."""
    out_bc, offs = okm.stride_coords(bc, stride)
    row = {tuple(c): i for i, c in enumerate(out_bc.tolist())}
    cells = bc.copy()
    cells[:, 1:] = np.floor_divide(bc[:, 1:], np.asarray(stride))
    owner = np.array([row[tuple(c)] for c in cells.tolist()])
    pooled = np.zeros((len(out_bc), x.shape[1]), np.float64)
    for m in range(len(out_bc)):
        sel = x[owner == m]
        pooled[m] = {"max": sel.max(0), "mean": sel.mean(0), "sum": sel.sum(0),
                     "min": sel.min(0)}[how]
    return out_bc, offs, pooled


@pytest.mark.parametrize("how", ["max", "mean", "sum", "min"])
def test_sparse_reduce_matches_dictionary_pooling(how):
    from warpconvnet_b200.nn.functional.sparse_pool import sparse_reduce
    v, coords, feats = _voxels(cin=16, seed=3)
    pooled = sparse_reduce(v, 2, 2, reduction=how)
    bc = okm.batch_indexed([c.numpy() for c in coords])
    x = torch.cat(feats).double().numpy()
    out_bc, offs, ref = _pool_oracle(bc, x, (2, 2, 2), how)
    assert pooled.tensor_stride == (2, 2, 2)
    assert np.array_equal(pooled.batch_indexed_coordinates.cpu().numpy(), out_bc)
    assert pooled.offsets.tolist() == offs.tolist()
    assert np.abs(pooled.feature_tensor.double().cpu().numpy() - ref).max() < 1e-5
    # var goes through to_csr + row reductions like the reference
    var = sparse_reduce(v, 2, 2, reduction="var").feature_tensor.double().cpu().numpy()
    mean = _pool_oracle(bc, x, (2, 2, 2), "mean")[2]
    assert np.abs(var - (_pool_oracle(bc, x * x, (2, 2, 2), "mean")[2] - mean ** 2)).max() < 1e-4


def test_reduce_and_stride_mode():
    """stride_mode=REDUCE_AND_STRIDE = max-pool over the stride window, then the conv at stride 1
    on the pooled voxels (reference helper.py:275-288, 539-548)."""
    from warpconvnet_b200.nn.functional.sparse_conv import STRIDED_CONV_MODE
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(2)
    v, coords, feats = _voxels(cin=32, seed=5)
    v.batched_features.batched_tensor.requires_grad_(True)
    conv = SparseConv3d(32, 64, 3, stride=2, bias=False,
                        stride_mode=STRIDED_CONV_MODE.REDUCE_AND_STRIDE).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = conv(v)
    out.feature_tensor.float().sum().backward()
    bc = okm.batch_indexed([c.numpy() for c in coords])
    out_bc, offs, pooled = _pool_oracle(bc, torch.cat(feats).double().numpy(), (2, 2, 2), "max")
    assert out.tensor_stride == (2, 2, 2)
    assert np.array_equal(out.batch_indexed_coordinates.cpu().numpy(), out_bc)
    km = okm.generate_kernel_map(out_bc, out_bc, (1, 1, 1), (3, 3, 3))
    xp = torch.from_numpy(pooled).bfloat16().float()
    w = conv.weight.detach().cpu().bfloat16().float()
    ref = oconv.forward(xp, w, km["in_maps"], km["out_maps"], km["offsets"], len(out_bc))
    assert oconv.rel_max_err(out.feature_tensor, ref) < 1e-2
    g = v.batched_features.batched_tensor.grad
    assert g is not None and g.shape == (5500, 32) and bool(torch.isfinite(g).all())
    assert conv.weight.grad is not None and float(conv.weight.grad.abs().sum()) > 0


@pytest.mark.parametrize("cin,cout,groups", [(64, 64, 8), (128, 256, 4), (512, 512, 64), (32, 32, 2)])
def test_group_conv(cin, cout, groups):
    """Group conv vs per-group explicit oracle (reference tests/nn/test_sparse_conv.py:742-776,
    bar rdiff < 0.01)."""
    from warpconvnet_b200.nn.functional.sparse_conv import (sparse_conv_dgrad,
                                                            sparse_conv_forward, sparse_conv_wgrad)
    km, x, _, gy, n_in, n_out = _case(3000, cin, cout, torch.bfloat16)
    g = torch.Generator().manual_seed(5)
    w = (torch.randn(27, groups, cin // groups, cout // groups, generator=g)
         * (27 * cin // groups) ** -0.5).cuda().bfloat16()
    args = (km.in_maps.cpu().numpy(), km.out_maps.cpu().numpy(), km.offsets.numpy())
    ry = oconv.forward_grouped(x.float().cpu(), w.float().cpu(), *args, n_out)
    rdx, rdw = oconv.backward_grouped(gy.float().cpu(), x.float().cpu(), w.float().cpu(), *args)
    y = sparse_conv_forward(x, w, km, n_out, groups=groups)
    dx = sparse_conv_dgrad(gy, w, km, n_in, groups=groups)
    dw = sparse_conv_wgrad(x, gy, tuple(w.shape), km, groups=groups)
    assert oconv.rel_max_err(y, ry) < 1e-2 and oconv.rdiff(y, ry) < 1e-2
    assert oconv.rel_max_err(dx, rdx) < 1e-2
    assert oconv.rel_max_err(dw, rdw) < 1e-4


def test_cpu_tensors_are_rejected():
    from warpconvnet_b200 import _ops
    with pytest.raises(RuntimeError):
        _ops.hash_prepare(torch.zeros(16, dtype=torch.int64), torch.zeros(16, dtype=torch.int32))


@pytest.mark.parametrize("kind,n,stride,ks", [("S", 160, 1, 3), ("R", 30000, 1, 3), ("R", 40000, 2, 2),
                                              ("S", 448, 1, 3)])
@pytest.mark.parametrize("parts,rounds", [(2, 2), (4, 4), (8, 3), (1, 5)])
def test_wgrad_row_block_major_order(kind, n, stride, ks, parts, rounds):
    """The L2-locality unit order of wgrad (row blocks x offsets, round-robin chunks) contracts
    exactly the same pairs as the offset-major order: same dW up to fp32 summation order, and the
    oracle's dW on the small cases."""
    from oracle import conv as oconv
    from warpconvnet_b200 import _ops
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    c = surface_coords(n, 1) if kind == "S" else random_coords(n, 0.3, 1)
    m = len(c)
    bc = torch.from_numpy(np.concatenate([np.zeros((m, 1), np.int32), c], 1)).cuda()
    if stride == 1:
        out_bc = bc
    else:
        from warpconvnet_b200.geometry.coords.ops.stride import stride_coords
        out_bc, _ = stride_coords(bc, (stride,) * 3)
    km = generate_kernel_map(bc, out_bc, (stride,) * 3, (ks,) * 3, same_coords=stride == 1)
    K, cin, cout = ks ** 3, 64, 128
    g = torch.Generator().manual_seed(5)
    x = torch.randn(m, cin, generator=g).bfloat16().cuda()
    gy = torch.randn(len(out_bc), cout, generator=g).bfloat16().cuda()
    args = (x, gy, km._in_buf, km._out_buf, km.offsets_dev, K, 1, cin, cout)
    plain = _ops.wgrad(*args)
    bp = km._block_prefix
    assert bp.shape[0] == K
    if bp.shape[1] < parts:
        pytest.skip("fewer row blocks than parts")
    blocked = _ops.wgrad(*args, row_block_prefix=bp, row_parts=parts, rounds=rounds)
    torch.cuda.synchronize()
    scale = float(plain.abs().max())
    assert float((blocked - plain).abs().max()) <= 1e-4 * scale
    if stride == 1:
        # identity offset through TMA tiles (submanifold map, unique coordinates), alone and
        # combined with the row-block-major order; fp16 and odd channel counts too
        st = km._hashtable.status_tensor
        ident = _ops.wgrad(*args, identity_k=K // 2, status=st)
        both = _ops.wgrad(*args, row_block_prefix=bp, row_parts=parts, rounds=rounds,
                          identity_k=K // 2, status=st)
        assert float((ident - plain).abs().max()) <= 1e-4 * scale
        assert float((both - plain).abs().max()) <= 1e-4 * scale
        x2, g2 = x[:, :48].contiguous().half(), gy[:, :96].contiguous().half()
        a2 = (x2, g2, km._in_buf, km._out_buf, km.offsets_dev, K, 1, 48, 96)
        p2 = _ops.wgrad(*a2)
        i2 = _ops.wgrad(*a2, identity_k=K // 2, status=st)
        assert float((i2 - p2).abs().max()) <= 1e-4 * float(p2.abs().max())
    if m <= 60000:
        ref = oconv.backward(gy.float().cpu(), x.float().cpu(),
                             torch.zeros(K, cin, cout), km.in_maps.cpu().numpy(),
                             km.out_maps.cpu().numpy(), km.offsets.numpy())[1]
        assert oconv.rel_max_err(blocked.view(K, cin, cout), ref) < 1e-4


@pytest.mark.parametrize("groups,cin,cout", [(1, 64, 128), (1, 48, 16), (4, 32, 64)])
@pytest.mark.parametrize("src,dst", [(torch.float32, torch.bfloat16), (torch.float32, torch.float16),
                                     (torch.bfloat16, torch.bfloat16), (torch.float32, torch.float32)])
def test_weight_image_pair_equals_two_single_images(groups, cin, cout, src, dst):
    """wcn_weight_image_pair (one launch, optional fp32 -> 16-bit conversion) produces byte-identical
    images to cast + two wcn_weight_image calls."""
    from warpconvnet_b200 import _ops
    K = 27
    g = torch.Generator().manual_seed(3)
    w = torch.randn(K, groups, cin // groups, cout // groups, generator=g).to(src).cuda()
    cg, og = cin // groups, cout // groups
    a, b = _ops.weight_image_pair(w, K, groups, cg, og, dst)
    wc = w.to(dst).contiguous()
    assert torch.equal(a, _ops.weight_image(wc, K, groups, cg, og, transpose_w=False))
    assert torch.equal(b, _ops.weight_image(wc, K, groups, cg, og, transpose_w=True))
    only, none = _ops.weight_image_pair(w, K, groups, cg, og, dst, want_transposed=False)
    assert none is None and torch.equal(only, a)
