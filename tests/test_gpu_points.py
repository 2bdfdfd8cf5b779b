# SPDX-License-Identifier: Apache-2.0
"""GPU tests of the Points / PointConv row: device grid-kNN vs the brute-force
oracle, PointConv forward / backward vs the CPU restatement of the reference module."""
import copy

import numpy as np
import pytest
import torch

from oracle import points as opts

pytestmark = pytest.mark.gpu


def _cloud(sizes, seed=0, scale=(1.0, 1.0, 1.0)):
    g = torch.Generator().manual_seed(seed)
    pts = [torch.rand(n, 3, generator=g) * torch.tensor(scale) for n in sizes]
    offs = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int64)
    return torch.cat(pts), offs


@pytest.mark.parametrize("sizes,k", [((3000, 2000), 16), ((500,), 8), ((20000,), 16),
                                     ((70, 4000, 33), 32), ((100,), 64), ((5000, 5000), 1)])
def test_knn_vs_bruteforce(sizes, k):
    from warpconvnet_b200 import _ops
    ref, offs = _cloud(sizes, 0, scale=(1.0, 0.3, 2.0))  # anisotropic bounding box
    qry, qoffs = _cloud([max(n // 2, 1) for n in sizes], 1, scale=(1.0, 0.3, 2.0))
    idx, dist = _ops.knn_search(ref.cuda(), offs, qry.cuda(), qoffs, k, return_distances=True)
    ridx, rdist = opts.knn(ref.numpy(), offs.numpy(), qry.numpy(), qoffs.numpy(), k)
    assert idx.shape == (qry.shape[0], k) and idx.dtype == torch.int64
    # distances are the invariant (fp32 kernel vs fp64 oracle); indices agree wherever the
    # neighbouring distances are separated by more than fp32 noise
    assert np.allclose(dist.cpu().numpy(), rdist, rtol=1e-5, atol=1e-6)
    got = idx.cpu().numpy()
    gap_ok = np.ones_like(rdist, dtype=bool)
    gap_ok[:, 1:] &= (rdist[:, 1:] - rdist[:, :-1]) > 1e-5
    gap_ok[:, :-1] &= (rdist[:, 1:] - rdist[:, :-1]) > 1e-5
    assert np.array_equal(got[gap_ok], ridx[gap_ok])
    # neighbours stay inside the query's batch item
    b_of = np.searchsorted(offs.numpy(), got, side="right") - 1
    q_b = np.searchsorted(qoffs.numpy(), np.arange(qry.shape[0]), side="right") - 1
    assert np.array_equal(b_of, np.repeat(q_b[:, None], k, 1))


def test_knn_self_query_and_duplicates():
    from warpconvnet_b200 import _ops
    ref, offs = _cloud((1000,), 3)
    ref[10] = ref[11]  # exact duplicate: smaller index first
    idx, dist = _ops.knn_search(ref.cuda(), offs, ref.cuda(), offs, 4, return_distances=True)
    assert torch.equal(idx[:, 0].cpu()[:10], torch.arange(10))
    assert idx[11, 0].item() == 10 and idx[11, 1].item() == 11 and idx[10, 0].item() == 10
    assert float(dist[:, 0].max()) == 0.0


def test_knn_rejects_bad_arguments():
    from warpconvnet_b200.geometry.coords.search.knn import batched_knn_search
    ref, offs = _cloud((10,), 0)
    with pytest.raises(AssertionError):
        batched_knn_search(ref.cuda(), offs, ref.cuda(), offs, 16)
    with pytest.raises(RuntimeError):
        batched_knn_search(ref, offs, ref, offs, 4)


@pytest.mark.parametrize("use_rel_pos,reductions", [(False, ("mean",)), (True, ("mean", "max")),
                                                    (False, ("sum", "min"))])
def test_point_conv_forward_backward(use_rel_pos, reductions):
    from warpconvnet_b200.geometry.coords.search.search_configs import RealSearchConfig
    from warpconvnet_b200.geometry.types.points import Points
    from warpconvnet_b200.nn.modules.point_conv import PointConv
    torch.manual_seed(0)
    sizes, cin, cout, k = (1500, 1200), 16, 32, 16
    coords, offs = _cloud(sizes, 5)
    feats = torch.randn(sum(sizes), cin)
    conv = PointConv(cin, cout, RealSearchConfig("knn", knn_k=k), use_rel_pos=use_rel_pos,
                     reductions=reductions).cuda()
    pc = Points(coords.cuda(), feats.cuda().requires_grad_(True), offsets=offs)
    out = pc_out = conv(pc)
    assert isinstance(out, Points) and out.feature_tensor.shape == (sum(sizes), cout)
    g = torch.randn_like(out.feature_tensor)
    out.feature_tensor.backward(g)
    # CPU fp64 restatement of the reference module on the oracle's neighbours
    ridx, _ = opts.knn(coords.numpy(), offs.numpy(), coords.numpy(), offs.numpy(), k)
    ref_conv = copy.deepcopy(conv).cpu().double()
    xin = feats.double().requires_grad_(True)
    ref_out = opts.point_conv_forward(xin, xin, coords.double(), coords.double(), ridx, k,
                                      ref_conv.edge_transform_mlp, ref_conv.out_transform_mlp,
                                      reductions=reductions, use_rel_pos=use_rel_pos)
    ref_out.backward(g.cpu().double())
    err = (out.feature_tensor.detach().cpu().double() - ref_out.detach()).abs().max() \
        / ref_out.detach().abs().max()
    assert float(err) < 2e-4
    gx = pc.batched_features.batched_tensor.grad.cpu().double()
    assert float((gx - xin.grad).abs().max() / xin.grad.abs().max()) < 2e-3
    w = conv.edge_transform_mlp.block[0].weight.grad.cpu().double()
    rw = ref_conv.edge_transform_mlp.block[0].weight.grad
    assert float((w - rw).abs().max() / rw.abs().max()) < 2e-3
    # neighbour search is cached on the Points object (points.py:237-272)
    assert pc.neighbors(conv.neighbor_search_args) is pc.neighbors(conv.neighbor_search_args)
    assert pc_out.offsets.tolist() == offs.tolist()


def test_point_conv_provided_queries_bf16():
    from warpconvnet_b200.geometry.coords.search.search_configs import RealSearchConfig
    from warpconvnet_b200.geometry.types.points import Points
    from warpconvnet_b200.nn.modules.point_conv import PointConv
    torch.manual_seed(1)
    coords, offs = _cloud((4000,), 7)
    qcoords, qoffs = _cloud((900,), 8)
    conv = PointConv(64, 64, RealSearchConfig("knn", knn_k=16), out_point_type="provided",
                     provided_in_channels=8).cuda()
    pc = Points(coords.cuda(), torch.randn(4000, 64).cuda(), offsets=offs)
    qc = Points(qcoords.cuda(), torch.randn(900, 8).cuda(), offsets=qoffs)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = conv(pc, qc)
    assert out.feature_tensor.shape == (900, 64)
    assert torch.isfinite(out.feature_tensor.float()).all()


def _sorted_rows(idx, dist, splits):
    idx, splits = np.asarray(idx), np.asarray(splits)
    dist = np.asarray(dist)
    out_i, out_d = idx.copy(), dist.copy()
    for q in range(len(splits) - 1):
        s, e = splits[q], splits[q + 1]
        order = np.argsort(idx[s:e], kind="stable")
        out_i[s:e], out_d[s:e] = idx[s:e][order], dist[s:e][order]
    return out_i, out_d


@pytest.mark.parametrize("name", ["radius_b2", "radius_b1_self"])
def test_radius_search_vs_reference_golden(name):
    import os
    from warpconvnet_b200.geometry.coords.search.radius import batched_radius_search
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    idx, dist, splits = batched_radius_search(
        torch.from_numpy(d["ref"]).cuda(), torch.from_numpy(d["ref_offsets"]),
        torch.from_numpy(d["query"]).cuda(), torch.from_numpy(d["query_offsets"]),
        float(d["radius"]))
    assert idx.dtype == torch.int64 and splits.dtype == torch.int64
    assert np.array_equal(splits.cpu().numpy(), d["splits"])
    gi, gd = _sorted_rows(d["idx"], d["dist"], d["splits"])
    oi, od = _sorted_rows(idx.cpu().numpy(), dist.cpu().numpy(), d["splits"])
    assert np.array_equal(oi, gi)
    assert np.allclose(od, gd, atol=1e-3)   # fp32 matmul-cdist error of the reference (up to 7e-4 at d = 0)


@pytest.mark.parametrize("sizes,qsizes,r", [((6000, 4000), (2500, 1500), 0.05),
                                             ((5000,), (200,), 0.5), ((300, 1, 800), (50, 7, 60), 0.2)])
def test_radius_search_vs_oracle(sizes, qsizes, r):
    from oracle import points as opts
    from warpconvnet_b200.geometry.coords.search.radius import batched_radius_search
    g = torch.Generator().manual_seed(len(sizes) * 7 + 1)
    ref = torch.rand(sum(sizes), 3, generator=g)
    query = torch.rand(sum(qsizes), 3, generator=g) * 1.2 - 0.1   # some queries outside the bbox
    ro = np.concatenate([[0], np.cumsum(sizes)])
    qo = np.concatenate([[0], np.cumsum(qsizes)])
    idx, dist, splits = batched_radius_search(ref.cuda(), torch.from_numpy(ro), query.cuda(),
                                              torch.from_numpy(qo), r)
    # the oracle works in float64 on the float32 inputs; pairs whose distance is within 1e-6 of
    # the radius may legitimately differ, so compare through a tolerance band
    oi, od, osplit = opts.radius(ref.numpy(), ro, query.numpy(), qo, r)
    lo_i, _, lo_s = opts.radius(ref.numpy(), ro, query.numpy(), qo, r - 1e-5)
    hi_i, _, hi_s = opts.radius(ref.numpy(), ro, query.numpy(), qo, r + 1e-5)
    gi = idx.cpu().numpy()
    gs = splits.cpu().numpy()
    assert len(gs) == len(osplit)
    for q in range(len(gs) - 1):
        got = set(gi[gs[q]:gs[q + 1]].tolist())
        assert set(lo_i[lo_s[q]:lo_s[q + 1]].tolist()) <= got <= set(hi_i[hi_s[q]:hi_s[q + 1]].tolist())
        assert len(got) == gs[q + 1] - gs[q]            # no duplicates
    gd = dist.cpu().numpy()
    dd = np.linalg.norm(ref.numpy()[gi].astype(np.float64)
                        - np.repeat(query.numpy().astype(np.float64), np.diff(gs), axis=0), axis=1)
    assert np.allclose(gd, dd, atol=1e-5)


def test_point_conv_radius_mode_runs():
    from warpconvnet_b200.geometry.coords.search.search_configs import RealSearchConfig
    from warpconvnet_b200.geometry.types.points import Points
    from warpconvnet_b200.nn.modules.point_conv import PointConv
    torch.manual_seed(0)
    coords = [torch.rand(2000, 3), torch.rand(1500, 3)]
    feats = [torch.randn(2000, 16), torch.randn(1500, 16)]
    pc = Points(coords, feats, device="cuda")
    conv = PointConv(16, 32, neighbor_search_args=RealSearchConfig("radius", radius=0.1)).cuda()
    out = conv(pc)
    assert out.feature_tensor.shape == (3500, 32)
    out.feature_tensor.sum().backward()
    assert all(p.grad is not None for p in conv.parameters())


@pytest.mark.parametrize("reduction", ["random", "mean", "sum", "max", "min"])
def test_points_to_voxels_vs_numpy(reduction):
    """Points.to_voxels: quantise, unique per batch item, reduce features (reference:
    geometry/types/conversion/to_voxels.py:4-24); checked against a NumPy dictionary build."""
    from warpconvnet_b200.geometry.types.conversion.to_voxels import points_to_voxels
    from warpconvnet_b200.geometry.types.points import Points
    g = torch.Generator().manual_seed(9)
    coords = [torch.rand(3000, 3, generator=g) * 2 - 0.7, torch.rand(1200, 3, generator=g)]
    feats = [torch.randn(3000, 5, generator=g), torch.randn(1200, 5, generator=g)]
    pc = Points(coords, feats, device="cuda")
    pc.batched_features.batched_tensor.requires_grad_(True)
    vs = 0.1
    vox, inv = points_to_voxels(pc, vs, reduction, return_to_unique=True)
    # oracle
    cells, rows = {}, []
    for b, (c, f) in enumerate(zip(coords, feats)):
        q = np.floor(c.numpy() / np.float32(vs)).astype(np.int64)
        for i in range(len(q)):
            cells.setdefault((b, *q[i].tolist()), []).append(f[i].numpy())
    keys = sorted(cells)
    red = {"random": lambda a: a[0], "mean": lambda a: np.mean(a, 0), "sum": lambda a: np.sum(a, 0),
           "max": lambda a: np.max(a, 0), "min": lambda a: np.min(a, 0)}[reduction]
    ref_feats = np.stack([red(np.stack(cells[k])) for k in keys])
    ref_bc = np.array(keys, np.int64)
    assert np.array_equal(vox.batch_indexed_coordinates.cpu().numpy(), ref_bc)
    assert np.allclose(vox.feature_tensor.detach().cpu().numpy(), ref_feats, atol=1e-5)
    counts = np.bincount(ref_bc[:, 0], minlength=2)
    assert vox.offsets.tolist() == [0, counts[0], counts[0] + counts[1]]
    assert inv.shape == (4200,) and int(inv.max()) == len(keys) - 1
    vox.feature_tensor.sum().backward()           # features stay differentiable
    assert pc.batched_features.batched_tensor.grad is not None
    # the result feeds a sparse conv
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    out = SparseConv3d(5, 16, 3).cuda()(pc.to_voxels(vs))
    assert out.feature_tensor.shape == (len(keys), 16)
