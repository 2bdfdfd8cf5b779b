# SPDX-License-Identifier: Apache-2.0
"""The remaining constructor modes of ``SparseConv3d`` / ``SparseConv2d`` against the CPU oracle:
generative output coordinates (reference helper.py:58-144: expand, stride-then-expand,
scale-then-expand for transposed), 2-D convolution (torch_discrete.py:328-342: z padded with 0),
anisotropic kernel size / stride / dilation, and a seeded sweep of kernel-map shapes.
Kernel maps bit-exact; features vs the fp64 explicit oracle on the same operands."""
import numpy as np
import pytest
import torch

from conftest import random_coords
from oracle import conv as oconv
from oracle import kernel_map as okm

pytestmark = pytest.mark.gpu


def _expand_oracle(bc, ks, dilation=None):
    """Union of coord + offset_k over the kernel, unique, sorted by (batch, x, y, z)
    (geometry/coords/ops/expand.py:17-75)."""
    offs = okm.kernel_offsets(ks, dilation)
    allc = np.repeat(bc[None, :, :], len(offs), 0).astype(np.int64)
    allc[:, :, 1:] += offs[:, None, :]
    out = np.unique(allc.reshape(-1, 4), axis=0).astype(np.int32)
    nb = int(bc[:, 0].max()) + 1
    offsets = np.zeros(nb + 1, np.int64)
    np.cumsum(np.bincount(out[:, 0], minlength=nb), out=offsets[1:])
    return out, offsets


def _scenes(sizes, cin, seed=0, scale=1, tensor_stride=None):
    from warpconvnet_b200.geometry.types.voxels import Voxels
    g = torch.Generator().manual_seed(seed)
    coords = [torch.from_numpy(random_coords(n, 0.3, seed + i) * scale) for i, n in enumerate(sizes)]
    feats = [torch.randn(n, cin, generator=g) for n in sizes]
    kw = {} if tensor_stride is None else {"tensor_stride": tensor_stride}
    v = Voxels(coords, feats, device="cuda", **kw)
    v.batched_features.batched_tensor.requires_grad_(True)
    return v, okm.batch_indexed([c.numpy() for c in coords]), torch.cat(feats)


def _check(conv, v, out, x, in_maps, out_maps, offsets, out_bc, tol=5e-3):
    assert np.array_equal(out.batch_indexed_coordinates.cpu().numpy(), out_bc)
    w = conv.weight.detach().cpu()
    ref = oconv.forward(x, w, in_maps, out_maps, offsets, len(out_bc))
    assert oconv.rel_max_err(out.feature_tensor, ref) < tol
    g = torch.Generator().manual_seed(9)
    gy = torch.randn(len(out_bc), w.shape[-1], generator=g)
    out.feature_tensor.backward(gy.cuda())
    dx_ref, dw_ref = oconv.backward(gy, x, w, in_maps, out_maps, offsets)
    assert oconv.rel_max_err(v.batched_features.batched_tensor.grad, dx_ref) < tol
    assert oconv.rel_max_err(conv.weight.grad, dw_ref) < 1e-3


def test_generative_stride1_expands_coordinates():
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(0)
    v, bc, x = _scenes((1500, 900), 16)
    conv = SparseConv3d(16, 32, 3, generative=True, bias=False).cuda()
    out = conv(v)
    out_bc, offs = _expand_oracle(bc, (3, 3, 3))
    assert out.offsets.tolist() == offs.tolist() and len(out_bc) > len(bc)
    km = okm.generate_kernel_map(bc, out_bc, (1, 1, 1), (3, 3, 3))
    _check(conv, v, out, x, km["in_maps"], km["out_maps"], km["offsets"], out_bc)


def test_generative_strided():
    """stride 2, kernel 3, generative: stride the coordinates, expand them, map the ORIGINAL
    input onto the expanded set with stride 2 (helper.py:122-144, STRIDE_ONLY)."""
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(1)
    v, bc, x = _scenes((2000, 1200), 16, seed=2)
    conv = SparseConv3d(16, 32, 3, stride=2, generative=True, bias=False).cuda()
    out = conv(v)
    strided, _ = okm.stride_coords(bc, (2, 2, 2))
    out_bc, offs = _expand_oracle(strided, (3, 3, 3))
    assert out.tensor_stride == (2, 2, 2) and out.offsets.tolist() == offs.tolist()
    km = okm.generate_kernel_map(bc, out_bc, (2, 2, 2), (3, 3, 3))
    _check(conv, v, out, x, km["in_maps"], km["out_maps"], km["offsets"], out_bc)


def test_generative_transposed_upsamples():
    """transposed + generative, stride 2, kernel 2: input coordinates are scaled by the stride,
    expanded by the kernel, and the map is built out -> in at stride 1 and swapped
    (helper.py:101-120, 513-530)."""
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(2)
    v, bc, x = _scenes((800, 500), 32, seed=4, tensor_stride=2)
    conv = SparseConv3d(32, 16, 2, stride=2, transposed=True, generative=True, bias=False).cuda()
    out = conv(v)
    scaled = bc.copy()
    scaled[:, 1:] *= 2
    out_bc, offs = _expand_oracle(scaled, (2, 2, 2))
    assert out.tensor_stride == (1, 1, 1) and out.offsets.tolist() == offs.tolist()
    km = okm.generate_kernel_map(out_bc, scaled, (1, 1, 1), (2, 2, 2))
    # every scaled input voxel reaches its 2^3 children
    assert int(km["offsets"][-1]) == 8 * len(bc)
    _check(conv, v, out, x, km["out_maps"], km["in_maps"], km["offsets"], out_bc)


@pytest.mark.parametrize("ks,stride", [(3, 1), (2, 2), ((3, 5), 1)])
def test_sparse_conv2d(ks, stride):
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv2d
    torch.manual_seed(3)
    g = torch.Generator().manual_seed(5)
    coords = []
    for n, side in ((1800, 64), (700, 40)):
        idx = torch.randperm(side * side, generator=g)[:n]
        coords.append(torch.stack([idx // side, idx % side], 1).int())
    feats = [torch.randn(len(c), 24, generator=g) for c in coords]
    v = Voxels(coords, feats, device="cuda")
    v.batched_features.batched_tensor.requires_grad_(True)
    conv = SparseConv2d(24, 40, ks, stride=stride, bias=False).cuda()
    out = conv(v)
    bc = okm.batch_indexed([c.numpy() for c in coords])
    ks2 = (ks, ks) if isinstance(ks, int) else ks
    st2 = (stride, stride)
    out_bc = bc if stride == 1 else okm.stride_coords(bc, st2)[0]
    km = okm.generate_kernel_map(bc, out_bc, st2, ks2)
    assert conv.weight.shape == (ks2[0] * ks2[1], 24, 40)
    _check(conv, v, out, torch.cat(feats), km["in_maps"], km["out_maps"], km["offsets"], out_bc)


def test_anisotropic_kernel_stride_dilation_module():
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(4)
    v, bc, x = _scenes((2500, 1500), 16, seed=6)
    conv = SparseConv3d(16, 24, (3, 1, 5), stride=(2, 1, 2), dilation=(1, 1, 2), bias=False).cuda()
    out = conv(v)
    out_bc, _ = okm.stride_coords(bc, (2, 1, 2))
    km = okm.generate_kernel_map(bc, out_bc, (2, 1, 2), (3, 1, 5), dilation=(1, 1, 2))
    assert out.tensor_stride == (2, 1, 2) and conv.weight.shape[0] == 15
    _check(conv, v, out, x, km["in_maps"], km["out_maps"], km["offsets"], out_bc)


def test_kernel_map_shape_sweep():
    """Seeded sweep over kernel size / stride / dilation / batch layout, negative coordinates
    included: offsets, pair table and CSR lists bit-exact vs the oracle for every draw."""
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    rng = np.random.RandomState(1234)
    for it in range(24):
        ks = tuple(int(k) for k in rng.choice([1, 2, 3, 4, 5], size=3))
        if int(np.prod(ks)) > 64:
            ks = (3, 3, int(ks[2] if ks[2] <= 5 else 3))
        stride = tuple(int(s) for s in rng.choice([1, 1, 2, 3], size=3))
        dil = tuple(int(d) for d in rng.choice([1, 1, 2], size=3))
        sizes = [int(s) for s in rng.randint(1, 3000, size=rng.randint(1, 4))]
        shift = rng.randint(-40, 5, size=3).astype(np.int32)
        bc = okm.batch_indexed([random_coords(n, float(rng.uniform(0.05, 0.5)), it * 7 + j) + shift
                                for j, n in enumerate(sizes)])
        out_bc = bc if stride == (1, 1, 1) else okm.stride_coords(bc, stride)[0]
        ti = torch.from_numpy(np.ascontiguousarray(bc)).cuda()
        to = ti if out_bc is bc else torch.from_numpy(np.ascontiguousarray(out_bc)).cuda()
        km = generate_kernel_map(ti, to, stride, ks, kernel_dilation=dil)
        ref = okm.generate_kernel_map(bc, out_bc, stride, ks, dilation=dil)
        tag = f"draw {it}: ks={ks} stride={stride} dil={dil} sizes={sizes}"
        assert np.array_equal(km.offsets.numpy(), ref["offsets"]), tag
        assert np.array_equal(km._pair_table.cpu().numpy(), ref["pair_table"]), tag
        assert np.array_equal(km.in_maps.cpu().numpy(), ref["in_maps"]), tag
        assert np.array_equal(km.out_maps.cpu().numpy(), ref["out_maps"]), tag
        iden = ref["identity_map_index"]
        assert km.identity_map_index == iden, tag


def test_explicit_output_coordinates_non_transposed():
    """``forward(x, output_spatially_sparse_tensor=target)``: the conv is evaluated at the
    target's coordinates (helper.py:421-428); here a 3^3 stride-1 conv sampled on another set."""
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(5)
    v, bc, x = _scenes((2000, 1000), 16, seed=8)
    tcoords = [torch.from_numpy(random_coords(n, 0.3, 20 + i)) for i, n in enumerate((1500, 1200))]
    target = Voxels(tcoords, [torch.zeros(len(c), 1) for c in tcoords], device="cuda")
    conv = SparseConv3d(16, 32, 3, bias=False).cuda()
    out = conv(v, target)
    out_bc = okm.batch_indexed([c.numpy() for c in tcoords])
    km = okm.generate_kernel_map(bc, out_bc, (1, 1, 1), (3, 3, 3))
    assert out.offsets.tolist() == target.offsets.tolist()
    _check(conv, v, out, x, km["in_maps"], km["out_maps"], km["offsets"], out_bc)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_compute_dtype_argument_and_grouped_strided(dtype):
    """``compute_dtype`` without autocast (fp32 features cast on entry, helper.py:310-320) on a
    grouped, strided conv; oracle on the rounded operands, per-group loops."""
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(6)
    v, bc, x = _scenes((3000,), 32, seed=10)
    conv = SparseConv3d(32, 64, 2, stride=2, groups=4, bias=False, compute_dtype=dtype).cuda()
    out = conv(v)
    assert out.feature_tensor.dtype == dtype and conv.weight.shape == (8, 4, 8, 16)
    out_bc, _ = okm.stride_coords(bc, (2, 2, 2))
    km = okm.generate_kernel_map(bc, out_bc, (2, 2, 2), (2, 2, 2))
    xr = x.to(dtype).float()
    wr = conv.weight.detach().cpu().to(dtype).float()
    ref = oconv.forward_grouped(xr, wr, km["in_maps"], km["out_maps"], km["offsets"], len(out_bc))
    assert np.array_equal(out.batch_indexed_coordinates.cpu().numpy(), out_bc)
    assert oconv.rel_max_err(out.feature_tensor, ref) < 1e-2
    g = torch.Generator().manual_seed(11)
    gy = torch.randn(len(out_bc), 64, generator=g).to(dtype)
    out.feature_tensor.backward(gy.cuda())
    dx_ref, dw_ref = oconv.backward_grouped(gy.float(), xr, wr, km["in_maps"], km["out_maps"],
                                            km["offsets"])
    assert oconv.rel_max_err(v.batched_features.batched_tensor.grad, dx_ref) < 1e-2
    assert oconv.rel_max_err(conv.weight.grad, dw_ref) < 1e-3


def test_tile_plan_is_a_row_permutation_and_close_to_the_full_width_sort():
    """The 24-bit compressed mask keys (cuhash.cu, three radix passes instead of four) may order
    the rows differently from a numeric sort of the 27-bit masks; the plan must still list every
    row exactly once, give every tile exactly the offsets its rows use, and need at most 1 % more
    steps than the full-width order on surface data."""
    from conftest import surface_coords
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    bc = okm.batch_indexed([surface_coords(300, 1)])
    t = torch.from_numpy(bc).cuda()
    n = len(bc)
    km = generate_kernel_map(t, t, (1, 1, 1), (3, 3, 3), same_coords=True)
    plan = km.fwd_plan(n)
    rows = plan.rows.cpu().numpy()
    assert np.array_equal(np.sort(rows[rows >= 0]), np.arange(n))
    pt = km._pair_table.cpu().numpy()
    masks = np.zeros(n, np.uint32)
    for k in range(27):
        masks |= (pt[k] >= 0).astype(np.uint32) << np.uint32(k)
    tr = plan.tile_rows

    def union_steps(order):
        m = np.concatenate([masks[order], np.zeros((-n) % tr, np.uint32)]).reshape(-1, tr)
        return np.array([bin(int(u)).count("1") for u in np.bitwise_or.reduce(m, axis=1)])

    nk = plan.tile_nk.cpu().numpy()[:plan.num_tiles]
    padded = np.where(rows >= 0, rows, 0)
    mine = np.where(rows >= 0, masks[padded], 0).astype(np.uint32).reshape(-1, tr)
    assert np.array_equal(nk, [bin(int(u)).count("1") for u in np.bitwise_or.reduce(mine, axis=1)])
    full = union_steps(np.argsort(masks, kind="stable")).sum()
    assert nk.sum() <= 1.01 * full, (int(nk.sum()), int(full))
