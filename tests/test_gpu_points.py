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
