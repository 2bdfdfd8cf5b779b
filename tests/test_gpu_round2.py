# SPDX-License-Identifier: Apache-2.0
"""Round-2 additions, all through the C-ABI on the device:

* native coordinate-set chain (csrc/coords.cu): stride_coords / unique / expand_coords vs the oracle
  and the host implementation, edge cases (negative coordinates, empty batch items, duplicates,
  out-of-range rows), no implicit host sync on the conv path (torch sync-debug mode);
* whole-step CUDA-graph capture with the SizeTape (utils/graph.py) vs the eager step;
* BatchNorm fixes (running statistics in 16-bit dtypes, eval-mode affine gradients, n = 1);
* deferred kernel-map status: an out-of-range coordinate raises inside the step;
* the reference-dispatcher adapters (integration/reference_backend.py) executed on the GPU with
  synthetic FwdCtx / BwdCtx, and — when baseline/_ref holds the built reference — through the
  reference's own SparseConv3d with ``fwd_algo=["wcn_b200"]``, plus the kernel map cross-checked
  against the reference's ``_C.cuhash`` build.
"""
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import random_coords, surface_coords
from oracle import conv as oconv
from oracle import kernel_map as okm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
pytestmark = pytest.mark.gpu


def _bc(scenes):
    return okm.batch_indexed(scenes)


# ------------------------------------------------------------------------------------------------
# coordinate-set chain
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("stride", [(2, 2, 2), (2, 1, 2), (3, 3, 3), (4, 2, 1)])
def test_stride_coords_native_matches_oracle(stride):
    from warpconvnet_b200.geometry.coords.ops.stride import stride_coords
    scenes = [random_coords(3000, 0.3, 1) - 7, random_coords(1700, 0.2, 2) - 30,
              random_coords(2500, 0.4, 3) + 100]
    bc = _bc(scenes)
    out, offs = stride_coords(torch.from_numpy(bc).cuda(), stride, n_batches=3)
    ref, ref_offs = okm.stride_coords(bc, stride)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert offs.tolist() == [int(v) for v in ref_offs]


def test_stride_coords_empty_batch_items_and_unknown_batch_count():
    from warpconvnet_b200.geometry.coords.ops.stride import stride_coords
    a, b = random_coords(500, 0.3, 4), random_coords(400, 0.3, 5)
    bc = np.concatenate([np.concatenate([np.full((len(a), 1), 1, np.int32), a], 1),
                         np.concatenate([np.full((len(b), 1), 3, np.int32), b], 1)])
    out, offs = stride_coords(torch.from_numpy(bc).cuda(), (2, 2, 2), n_batches=5)
    ref, _ = okm.stride_coords(bc, (2, 2, 2))
    assert np.array_equal(out.cpu().numpy(), ref)
    n1 = int((ref[:, 0] == 1).sum())
    assert offs.tolist() == [0, 0, n1, n1, len(ref), len(ref)]
    out2, offs2 = stride_coords(torch.from_numpy(bc).cuda(), (2, 2, 2))   # batch count unknown
    assert np.array_equal(out2.cpu().numpy(), ref) and offs2.tolist() == [0, 0, n1, n1, len(ref)]


def test_unique_native_first_occurrence_and_offsets():
    from warpconvnet_b200.geometry.coords.ops.stride import unique_with_offsets
    base = _bc([random_coords(800, 0.3, 6) - 3, random_coords(600, 0.3, 7)])
    g = np.random.RandomState(0)
    dup = base[g.randint(0, len(base), 900)]
    allc = np.concatenate([base, dup])
    order = np.argsort(allc[:, 0], kind="stable")          # keep batch-sorted, duplicates anywhere
    allc = allc[order]
    t = torch.from_numpy(allc)
    u_gpu, i_gpu, o_gpu = unique_with_offsets(t.cuda(), 2)
    u_cpu, i_cpu, o_cpu = unique_with_offsets(t, 2)
    assert torch.equal(u_gpu.cpu(), u_cpu) and torch.equal(i_gpu.cpu(), i_cpu)
    assert o_gpu.tolist() == o_cpu.tolist()
    assert len(u_gpu) == len(base)


def test_expand_coords_native_matches_host_path():
    from warpconvnet_b200.geometry.coords.ops.expand import expand_coords
    bc = torch.from_numpy(_bc([random_coords(700, 0.2, 8) - 5, random_coords(300, 0.2, 9)]))
    for ks, dil in (((3, 3, 3), (1, 1, 1)), ((2, 2, 2), (1, 1, 1)), ((3, 1, 3), (2, 1, 1))):
        g_out, g_offs = expand_coords(bc.cuda(), ks, dil, n_batches=2)
        c_out, c_offs = expand_coords(bc, ks, dil, n_batches=2)
        assert torch.equal(g_out.cpu(), c_out) and g_offs.tolist() == c_offs.tolist()


def test_out_of_range_coordinates_raise_in_coordinate_chain_and_in_the_conv_step():
    from warpconvnet_b200.geometry.coords.ops.stride import stride_coords
    from warpconvnet_b200.geometry.coords.search.search_results import check_pending_kernel_maps
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    c = random_coords(500, 0.3, 10)
    c[17] = (5, 600000, 3)                                   # beyond the packed range, also after / 2
    bc = torch.from_numpy(_bc([c])).cuda()
    with pytest.raises(ValueError):
        stride_coords(bc, (2, 2, 2), n_batches=1)
    # submanifold conv: nothing on the conv path reads the kernel map on the host, the deferred
    # status must still surface before the step ends (polled in backward / at the next map build)
    conv = SparseConv3d(16, 16, 3, bias=False).cuda()
    v = Voxels([torch.from_numpy(c)], [torch.randn(len(c), 16)], device="cuda")
    v.batched_features.batched_tensor.requires_grad_(True)
    with pytest.raises(ValueError):
        out = conv(v)
        torch.cuda.synchronize()
        out.feature_tensor.sum().backward()
        check_pending_kernel_maps(block=True)
    check_pending_kernel_maps(block=True)                    # the list is drained


def test_no_implicit_host_sync_on_the_strided_conv_path():
    """torch's sync-debug mode turns every implicit synchronisation of the compute stream
    (.item(), .cpu(), nonzero, boolean-mask indexing, pageable H2D ...) into an error: a strided
    conv + transposed conv + submanifold conv step, forward and backward, must pass under it.
    (The coordinate chain waits on its own side-stream event, which is explicit and does not
    drain the compute stream.)"""
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(0)
    coords = [torch.from_numpy(surface_coords(80, s)).cuda() for s in (0, 1)]
    feats = [torch.randn(len(c), 16, device="cuda") for c in coords]
    down = SparseConv3d(16, 32, 2, 2, bias=False).cuda()
    mid = SparseConv3d(32, 32, 3, bias=False).cuda()
    up = SparseConv3d(32, 16, 2, 2, transposed=True, bias=False).cuda()

    def step():
        x = Voxels(coords, feats)
        x.batched_features.batched_tensor.requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = up(mid(down(x)), x)
        y.feature_tensor.float().square().mean().backward()

    step()                                                   # warm-up (lazy inits, offset tables)
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        step()
    finally:
        torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------
# whole-step CUDA graph
# ------------------------------------------------------------------------------------------------
def test_whole_network_step_captured_in_one_cuda_graph_matches_eager():
    from minkunet14 import MinkUNet14, surface_scene
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.utils.graph import capture_step
    torch.manual_seed(0)
    coords = [surface_scene(72, s).cuda() for s in (0, 1)]
    feats = [torch.randn(len(c), 3, device="cuda") for c in coords]
    net = MinkUNet14(3, 20).cuda()
    for m in net.modules():                                   # momentum-free statistics: replays
        if isinstance(m, torch.nn.BatchNorm1d):               # and eager passes stay comparable
            m.momentum = 0.0
    state = {}

    def step():
        x = Voxels(coords, feats)
        for p in net.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(x)
        loss = out.feature_tensor.float().square().mean()
        loss.backward()
        # detached copies only: a live autograd graph from the eager pass would keep AccumulateGrad
        # nodes bound to the eager stream while the capture runs on its own stream
        state["loss"] = loss.detach().clone()
        state["gsum"] = torch.stack([p.grad.float().abs().sum() for p in net.parameters()]).sum()
        return None

    step()
    torch.cuda.synchronize()
    eager = (float(state["loss"]), float(state["gsum"]))
    graph, tape, _ = capture_step(step, warmup=1)
    assert len(tape.entries) == 4                             # one coordinate set per level change
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    tape.verify()
    got = (float(state["loss"]), float(state["gsum"]))
    assert abs(got[0] - eager[0]) <= 1e-3 * abs(eager[0]) + 1e-6
    assert abs(got[1] - eager[1]) <= 2e-2 * abs(eager[1])     # bf16 atomics order in wgrad


def test_capture_without_a_size_tape_fails_loudly():
    from warpconvnet_b200.geometry.coords.ops.stride import stride_coords
    bc = torch.from_numpy(_bc([random_coords(400, 0.3, 11)])).cuda()
    stride_coords(bc, (2, 2, 2), n_batches=1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with pytest.raises(RuntimeError, match="SizeTape"):
        with torch.cuda.graph(g, capture_error_mode="relaxed"):
            stride_coords(bc, (2, 2, 2), n_batches=1)


# ------------------------------------------------------------------------------------------------
# BatchNorm fixes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_batchnorm_running_statistics_in_16bit_module(dtype):
    """model.half() / .bfloat16() turns the BatchNorm1d buffers into 2-byte types; the kernel
    updates fp32 temporaries (no out-of-bounds write) and the result matches torch."""
    from warpconvnet_b200.nn.modules.normalizations import BatchNorm
    torch.manual_seed(0)
    c = 48
    bn = BatchNorm(c).cuda().to(dtype)
    ref = torch.nn.BatchNorm1d(c).cuda().to(dtype)
    guard = torch.full((4096,), 7.0, device="cuda", dtype=dtype)  # neighbours in the allocator
    x = torch.randn(1000, c, device="cuda", dtype=dtype)
    y = bn(x)
    y_ref = ref(x)
    torch.cuda.synchronize()
    assert bn.norm.running_mean.dtype == dtype
    assert torch.allclose(bn.norm.running_mean.float(), ref.running_mean.float(), atol=2e-2)
    assert torch.allclose(bn.norm.running_var.float(), ref.running_var.float(), atol=2e-2)
    assert torch.allclose(y.float(), y_ref.float(), atol=6e-2)
    assert bool((guard == 7.0).all())


def test_batchnorm_eval_mode_affine_gradients_match_torch():
    from warpconvnet_b200.nn.functional.normalizations import batch_norm_act
    torch.manual_seed(1)
    n, c = 700, 32
    x = torch.randn(n, c, device="cuda")
    rm, rv = torch.randn(c, device="cuda") * 0.1, torch.rand(c, device="cuda") + 0.5
    w = torch.randn(c, device="cuda", requires_grad=True)
    b = torch.randn(c, device="cuda", requires_grad=True)
    xg = x.clone().requires_grad_(True)
    y = batch_norm_act(xg, w, b, rm, rv, training=False, relu=True)
    gy = torch.randn_like(y)
    y.backward(gy)
    w2, b2, x2 = (t.detach().double().requires_grad_(True) for t in (w, b, x))
    y2 = torch.relu(torch.nn.functional.batch_norm(x2, rm.double(), rv.double(), w2, b2, False))
    y2.backward(gy.double())
    assert torch.allclose(y.double(), y2, atol=1e-4)
    assert torch.allclose(w.grad.double(), w2.grad, rtol=1e-3, atol=1e-3)
    assert torch.allclose(b.grad.double(), b2.grad, rtol=1e-3, atol=1e-3)
    assert torch.allclose(xg.grad.double(), x2.grad, rtol=1e-3, atol=1e-4)


def test_batchnorm_single_row_in_training_raises_and_bad_buffers_are_rejected():
    from warpconvnet_b200 import _ops
    from warpconvnet_b200._lib import WcnError
    from warpconvnet_b200.nn.functional.normalizations import batch_norm_act
    x = torch.randn(1, 8, device="cuda")
    with pytest.raises(ValueError):
        batch_norm_act(x, None, None, None, None, training=True)
    x = torch.randn(64, 8, device="cuda")
    with pytest.raises(WcnError):   # a 2-byte buffer must never reach the float* parameter
        _ops.bn_forward(x, None, None, 1e-5, 0.1, torch.zeros(8, device="cuda", dtype=torch.bfloat16),
                        torch.ones(8, device="cuda"), None, False)


# ------------------------------------------------------------------------------------------------
# reference-dispatcher adapters on the GPU
# ------------------------------------------------------------------------------------------------
def _oracle_case(n=3000, cin=32, cout=64, seed=0):
    c = random_coords(n, 0.3, seed)
    bc = _bc([c])
    km = okm.generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3))
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, generator=g)
    w = torch.randn(27, cin, cout, generator=g) * (27 * cin) ** -0.5
    gy = torch.randn(n, cout, generator=g)
    return bc, km, x, w, gy


def test_reference_backend_adapters_execute_on_gpu_with_synthetic_ctx():
    from warpconvnet_b200.integration import reference_backend as rb
    bc, km, x, w, gy = _oracle_case()
    n = len(bc)
    # what the reference hands its backends: CSR maps on the device, offsets on the CPU
    ref_map = types.SimpleNamespace(in_maps=torch.from_numpy(km["in_maps"]).cuda(),
                                    out_maps=torch.from_numpy(km["out_maps"]).cuda(),
                                    offsets=torch.from_numpy(km["offsets"]).long(),
                                    identity_map_index=13)
    fctx = types.SimpleNamespace(in_features=x.cuda(), weight=w.cuda(), kernel_map=ref_map,
                                 num_out_coords=n, compute_dtype=torch.bfloat16, groups=1)
    y = rb.forward_adapter(fctx)
    assert isinstance(y, torch.Tensor) and y.dtype == torch.float32
    xb, wb, gb = x.bfloat16().float(), w.bfloat16().float(), gy.bfloat16().float()
    args = (km["in_maps"], km["out_maps"], km["offsets"])
    assert oconv.rel_max_err(y, oconv.forward(xb, wb, *args, n)) < 1e-2
    bctx = types.SimpleNamespace(grad_output=gy.cuda(), in_features=x.cuda(), weight=w.cuda(),
                                 kernel_map=ref_map, compute_dtype=torch.bfloat16, groups=1,
                                 needs_input_grad=(True, True))
    dx, dw = rb.backward_adapter(bctx)
    dx_ref, dw_ref = oconv.backward(gb, xb, wb, *args)
    assert oconv.rel_max_err(dx, dx_ref) < 1e-2 and oconv.rel_max_err(dw, dw_ref) < 1e-2
    # unsupported shape -> negative status, not an exception (backends.py:489-510)
    bad = types.SimpleNamespace(in_features=x.cuda(), weight=torch.randn(27, 4, 3, 5).cuda(),
                                kernel_map=ref_map, num_out_coords=n,
                                compute_dtype=torch.bfloat16, groups=4)
    assert rb.forward_adapter(bad) == rb.STATUS_UNSUPPORTED


def _load_built_reference():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "warpconvnet")):
        pytest.skip("baseline/_ref (the built reference) is not present")
    os.environ.setdefault("WARPCONVNET_BENCHMARK_CACHE_DIR", os.path.join(ROOT, "gpurun_out", "ref_cache"))
    os.environ.setdefault("WARPCONVNET_AUTOTUNE_LOG", "false")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import warpconvnet  # noqa: F401
    except Exception as exc:
        pytest.skip(f"built reference does not import: {type(exc).__name__}: {exc}")


def test_kernel_map_equals_the_reference_cuhash_build():
    """offsets equal and per-offset pair SETS equal against the reference's own _C.cuhash kernel
    map on the same coordinates (stride 1 / 3^3 and stride 2 / 2^3)."""
    _load_built_reference()
    from warpconvnet.geometry.coords.search.torch_discrete import generate_kernel_map as ref_gkm
    from warpconvnet_b200.geometry.coords.ops.stride import stride_coords
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    for c, ks, st in ((surface_coords(200, 0), 3, 1), (random_coords(40000, 0.3, 1), 3, 1),
                      (surface_coords(160, 2), 2, 2), (random_coords(20000, 0.3, 3) - 40, 3, 2)):
        bc = torch.from_numpy(_bc([c])).cuda()
        out_bc = bc if st == 1 else stride_coords(bc, (st,) * 3, n_batches=1)[0]
        rk = ref_gkm(bc, out_bc, (st,) * 3, (ks,) * 3)
        ok_ = generate_kernel_map(bc, out_bc, (st,) * 3, (ks,) * 3)
        assert torch.equal(rk.offsets.cpu().long(), ok_.offsets.cpu().long())
        assert rk.identity_map_index == ok_.identity_map_index
        n_in = len(bc)
        for k in range(len(rk)):
            (ri, ro), (oi, oo) = rk[k], ok_[k]
            a = torch.sort(ro.long() * n_in + ri.long()).values
            b = torch.sort(oo.long() * n_in + oi.long()).values
            assert torch.equal(a, b), f"pair set of offset {k} differs"


def test_reference_sparseconv3d_runs_this_library_through_its_own_dispatcher():
    """INTEGRATION.md seam A end to end: the reference's SparseConv3d, Voxels and autograd
    function, with ``fwd_algo = dgrad_algo = wgrad_algo = ["wcn_b200"]``, against the oracle."""
    _load_built_reference()
    from warpconvnet.geometry.types.voxels import Voxels as RVoxels
    from warpconvnet.nn.modules.sparse_conv import SparseConv3d as RConv
    from warpconvnet_b200.integration import reference_backend as rb
    name = rb.register()
    bc, km, x, w, gy = _oracle_case(n=5000, cin=64, cout=64, seed=3)
    n = len(bc)
    conv = RConv(64, 64, 3, bias=False, fwd_algo=[name], dgrad_algo=[name], wgrad_algo=[name]).cuda()
    with torch.no_grad():
        conv.weight.copy_(w.cuda())
    feats = x.cuda().bfloat16().requires_grad_(True)
    vox = RVoxels([torch.from_numpy(bc[:, 1:].copy()).cuda()], [feats])
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = conv(vox)
    out.feature_tensor.backward(gy.cuda().to(out.feature_tensor.dtype))
    torch.cuda.synchronize()
    xb, wb, gb = x.bfloat16().float(), w.bfloat16().float(), gy.bfloat16().float()
    args = (km["in_maps"], km["out_maps"], km["offsets"])
    assert oconv.rel_max_err(out.feature_tensor, oconv.forward(xb, wb, *args, n)) < 1e-2
    dx_ref, dw_ref = oconv.backward(gb, xb, wb, *args)
    assert oconv.rel_max_err(feats.grad, dx_ref) < 1e-2
    assert oconv.rel_max_err(conv.weight.grad, dw_ref) < 1e-2


# ------------------------------------------------------------------------------------------------
# PointConv values pinned by the reference's own module (tests/golden/make_golden_pointconv.py)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,cin,cout,k,kw", [
    ("pointconv_knn8", 16, 32, 8, {}),
    ("pointconv_knn16_relpos", 8, 24, 16, dict(use_rel_pos=True, reductions=("mean", "max"))),
])
def test_pointconv_matches_reference_generated_fixture(name, cin, cout, k, kw):
    """Same weights (the reference module's state_dict loads unchanged), same points: our
    PointConv (grid kNN kernel + CSR reductions) must reproduce the REFERENCE PointConv's output
    and neighbour lists (fp32, CPU reference: cdist + topk + segment_csr)."""
    from warpconvnet_b200.geometry.coords.search.search_configs import RealSearchConfig
    from warpconvnet_b200.geometry.types.points import Points
    from warpconvnet_b200.nn.modules.point_conv import PointConv
    d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    conv = PointConv(cin, cout, RealSearchConfig("knn", knn_k=k), **kw)
    state = {key: torch.from_numpy(d["p__" + key]) for key in d["state_keys"]}
    conv.load_state_dict(state, strict=True)
    conv = conv.cuda().eval()
    pc = Points(torch.from_numpy(d["coords"]).cuda(), torch.from_numpy(d["feats"]).cuda(),
                offsets=torch.from_numpy(d["offsets"]))
    with torch.no_grad():
        out = conv(pc)
    nbrs = pc.neighbors(RealSearchConfig("knn", knn_k=k))
    got_knn = nbrs.neighbor_indices.reshape(-1, k).cpu().numpy()
    # neighbour SETS per row (order among equidistant neighbours is unspecified upstream; the
    # random float coordinates of the fixture have no ties, so the ordered lists agree too)
    assert np.array_equal(np.sort(got_knn, 1), np.sort(d["knn"], 1))
    ref = torch.from_numpy(d["out"])
    assert out.feature_tensor.shape == ref.shape
    assert oconv.rel_max_err(out.feature_tensor, ref.double()) < 2e-4


def test_reference_sparseconv3d_through_the_capi_stub():
    """The self-contained ctypes binding of INTEGRATION.md seam A (integration/capi_backend_stub.py:
    ctypes + torch only, straight onto include/wcn_b200.h) registered in the reference's own
    dispatcher and driven through the reference's SparseConv3d, against the oracle."""
    _load_built_reference()
    import importlib.util
    from warpconvnet.geometry.types.voxels import Voxels as RVoxels
    from warpconvnet.nn.modules.sparse_conv import SparseConv3d as RConv
    spec = importlib.util.spec_from_file_location(
        "capi_backend_stub", os.path.join(ROOT, "warpconvnet_b200", "integration", "capi_backend_stub.py"))
    stub = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stub)
    name = stub.register(os.path.join(ROOT, "warpconvnet_b200", "csrc", "libwcn_b200.so"))
    bc, km, x, w, gy = _oracle_case(n=6000, cin=32, cout=64, seed=5)
    n = len(bc)
    conv = RConv(32, 64, 3, bias=False, fwd_algo=[name], dgrad_algo=[name], wgrad_algo=[name]).cuda()
    with torch.no_grad():
        conv.weight.copy_(w.cuda())
    feats = x.cuda().bfloat16().requires_grad_(True)
    vox = RVoxels([torch.from_numpy(bc[:, 1:].copy()).cuda()], [feats])
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = conv(vox)
    out.feature_tensor.backward(gy.cuda().to(out.feature_tensor.dtype))
    torch.cuda.synchronize()
    xb, wb, gb = x.bfloat16().float(), w.bfloat16().float(), gy.bfloat16().float()
    args = (km["in_maps"], km["out_maps"], km["offsets"])
    assert oconv.rel_max_err(out.feature_tensor, oconv.forward(xb, wb, *args, n)) < 1e-2
    dx_ref, dw_ref = oconv.backward(gb, xb, wb, *args)
    assert oconv.rel_max_err(feats.grad, dx_ref) < 1e-2
    assert oconv.rel_max_err(conv.weight.grad, dw_ref) < 1e-2


# ------------------------------------------------------------------------------------------------
# GEMM-epilogue fusion: bias, ReLU and the BatchNorm statistics
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("n,cin,cout,relu", [
    (40000, 128, 128, False),   # 256-row tiles, swap-AB form
    (40000, 64, 96, True),      # 256-row tiles, row-major form
    (40000, 32, 32, False),     # one 64-byte chunk per row
    (3000, 64, 128, True),      # 128-row tiles
    (3000, 48, 256, False),     # two 256-byte column chunks (16-bit) / four (fp32)
    (700, 16, 16, True),
])
def test_gemm_epilogue_statistics_match_the_stored_output(dtype, n, cin, cout, relu):
    """stats[0] = sum_r y, stats[1] = sum_r y^2 of the output AS STORED (bias, ReLU, rounding),
    for every kernel form; compared with an fp64 reduction of the stored tensor."""
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    from warpconvnet_b200.nn.functional.sparse_conv import sparse_conv_forward
    c = surface_coords(int(n ** 0.5), 1)
    bc = torch.from_numpy(_bc([c])).cuda()
    m = len(c)
    km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
    g = torch.Generator().manual_seed(n + cout)
    x = torch.randn(m, cin, generator=g).cuda().to(dtype)
    w = (torch.randn(27, cin, cout, generator=g) * (27 * cin) ** -0.5).cuda().to(dtype)
    bias = torch.randn(cout, generator=g).cuda()
    stats = torch.zeros((2, cout), dtype=torch.float64, device="cuda")
    y = sparse_conv_forward(x, w, km, m, bias=bias, relu=relu, stats=stats)
    y_plain = sparse_conv_forward(x, w, km, m, bias=bias, relu=relu)
    assert torch.equal(y, y_plain)                          # the statistics do not touch the output
    ref0 = y.double().sum(0)
    ref1 = y.double().square().sum(0)
    assert torch.allclose(stats[0], ref0, rtol=1e-5, atol=1e-4 * float(y.double().abs().sum(0).max()))
    assert torch.allclose(stats[1], ref1, rtol=1e-5, atol=1e-6 * float(ref1.max()))


def test_conv_bias_and_batchnorm_statistics_through_the_epilogue_module_path():
    """SparseConv3d(bias=True, emit_bn_stats) -> BatchNorm(relu): bias is added in the epilogue
    (and gets its gradient), BatchNorm consumes the epilogue's statistics; the result and every
    gradient equal the unfused path (separate statistics pass) and torch's BatchNorm."""
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules.normalizations import BatchNorm
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    import warpconvnet_b200._ops as ops
    torch.manual_seed(3)
    coords = [torch.from_numpy(surface_coords(90, s)) for s in (0, 1)]
    feats = [torch.randn(len(c), 32) for c in coords]
    results = {}
    for fused in (True, False):
        torch.manual_seed(4)
        conv = SparseConv3d(32, 64, 3, bias=True).cuda()
        conv.emit_bn_stats = fused
        bn = BatchNorm(64, relu=True).cuda()
        v = Voxels(coords, feats, device="cuda")
        v.batched_features.batched_tensor.requires_grad_(True)
        calls = []
        orig = ops.bn_forward
        ops.bn_forward = lambda *a, **k: (calls.append(k.get("sums") is not None), orig(*a, **k))[1]
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = bn(conv(v))
        finally:
            ops.bn_forward = orig
        assert calls == [fused]                               # the fused path really was taken
        out.feature_tensor.float().square().mean().backward()
        results[fused] = (out.feature_tensor.float(), conv.bias.grad.clone(), conv.weight.grad.clone(),
                          bn.norm.weight.grad.clone(), v.batched_features.batched_tensor.grad.clone(),
                          bn.norm.running_var.clone())
    for i, (a, b) in enumerate(zip(results[True], results[False])):
        if i == 1:
            # a bias in front of a BatchNorm has a zero gradient by construction (the norm removes
            # the channel mean): both paths must return rounding noise around 0
            scale = float(results[True][2].abs().max())
            assert float(a.abs().max()) < 2e-2 * scale and float(b.abs().max()) < 2e-2 * scale
            continue
        # the two paths add the same bf16 values in a different order: scale / shift differ in the
        # last fp32 bits, which can flip the bf16 rounding of single outputs (1 ulp = 3.9e-3)
        assert oconv.rel_max_err(a, b.double().cpu()) < 1e-2


def test_bias_through_the_epilogue_forward_and_gradient():
    """y = conv(x) + bias with the bias added in the GEMM epilogue; d bias = sum_r dY[r]."""
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(5)
    c = torch.from_numpy(surface_coords(70, 2))
    f = torch.randn(len(c), 32)
    conv = SparseConv3d(32, 48, 3, bias=True).cuda()
    v = Voxels([c], [f], device="cuda")
    out = conv(v)
    with torch.no_grad():
        saved = conv.bias.clone()
        conv.bias.zero_()
        base = conv(v).feature_tensor
        conv.bias.copy_(saved)
    assert torch.allclose(out.feature_tensor, base + saved, atol=2e-3)
    gy = torch.randn_like(out.feature_tensor)
    out.feature_tensor.backward(gy)
    assert torch.allclose(conv.bias.grad, gy.sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("groups,cin,cout", [(16, 64, 64), (8, 32, 16), (3, 12, 6)])
def test_group_conv_with_fewer_than_8_channels_per_group(groups, cin, cout):
    """The reference's mask_gemm path rejects groups of fewer than 8 channels
    (detail/dispatch.py:42-50); here they run on the dense kernels through the block-diagonal
    weight. Forward, dX and the [K, G, Cin/G, Cout/G] weight gradient vs the per-group oracle."""
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    torch.manual_seed(groups)
    c = random_coords(2500, 0.3, groups)
    bc = _bc([c])
    km = okm.generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3))
    f = torch.randn(len(c), cin)
    conv = SparseConv3d(cin, cout, 3, groups=groups, bias=False).cuda()
    assert conv.weight.shape == (27, groups, cin // groups, cout // groups)
    v = Voxels([torch.from_numpy(c)], [f], device="cuda")
    v.batched_features.batched_tensor.requires_grad_(True)
    out = conv(v)
    gy = torch.randn(len(c), cout)
    out.feature_tensor.backward(gy.cuda())
    w = conv.weight.detach().cpu()
    args = (km["in_maps"], km["out_maps"], km["offsets"])
    y_ref = oconv.forward_grouped(f, w, *args, len(c))
    dx_ref, dw_ref = oconv.backward_grouped(gy, f, w, *args)
    assert oconv.rel_max_err(out.feature_tensor, y_ref) < 5e-3      # fp32 in: TF32 tensor cores
    assert oconv.rel_max_err(v.batched_features.batched_tensor.grad, dx_ref) < 5e-3
    assert conv.weight.grad.shape == conv.weight.shape
    assert oconv.rel_max_err(conv.weight.grad, dw_ref) < 1e-3


# ------------------------------------------------------------------------------------------------
# single-kernel mask sort (csrc/cuhash.cu mask_sort_kernel) vs a stable argsort of the same keys
# ------------------------------------------------------------------------------------------------
def _narrowed(keys: torch.Tensor, K: int) -> torch.Tensor:
    """The sort key wcn_sort_rows_by_key derives from a K-bit mask (K <= 32): masks of 25..32 bits
    are compressed to 24 (centre bit of an odd K dropped, low bits XOR-folded)."""
    k = keys.clone() & 0xFFFFFFFF
    bits = K
    if K > 24:
        if K & 1:
            d = K // 2
            k = ((k >> (d + 1)) << d) | (k & ((1 << d) - 1))
            bits -= 1
        f = bits - 24
        if f > 0:
            k = (k >> f) ^ (k & ((1 << f) - 1))
    return k


@pytest.mark.parametrize("K", [8, 27, 32, 24, 17])
@pytest.mark.parametrize("M", [1, 31, 513, 7001, 200704, (1 << 20) + 5])
def test_mask_sort_is_a_stable_sort_of_the_narrowed_keys(K, M):
    from warpconvnet_b200 import _ops
    from warpconvnet_b200._lib import check, lib
    g = torch.Generator().manual_seed(K * 1000 + M % 997)
    # few distinct values (structured masks) mixed with random ones
    pool = torch.randint(0, 1 << K, (37,), generator=g, dtype=torch.int64)
    keys = torch.where(torch.rand(M, generator=g) < 0.7,
                       pool[torch.randint(0, 37, (M,), generator=g)],
                       torch.randint(0, 1 << K, (M,), generator=g, dtype=torch.int64)).cuda()
    rows = torch.empty(M, dtype=torch.int32, device="cuda")
    ws_bytes = lib.wcn_sort_workspace_bytes(M)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    for _ in range(2):  # twice on the same workspace: the barrier counter is re-armed per launch
        check(lib.wcn_sort_rows_by_key(keys.data_ptr(), M, K, rows.data_ptr(), ws.data_ptr(), ws_bytes,
                                       _ops._stream()), "sort_rows_by_key")
        torch.cuda.synchronize()
        expect = torch.argsort(_narrowed(keys, K), stable=True)
        assert torch.equal(rows.long(), expect)


@pytest.mark.parametrize("K,M", [(27, 200704), (27, 5000), (8, 30000), (9, 700), (32, 4097)])
def test_mask_sort_from_the_table_equals_the_sort_of_explicit_masks(K, M):
    """wcn_sort_rows_by_table derives the row masks inside the sort kernel: same permutation as
    wcn_mask_keys + wcn_sort_rows_by_key on the same table."""
    from warpconvnet_b200 import _ops
    from warpconvnet_b200._lib import check, lib
    g = torch.Generator().manual_seed(K + M)
    table = torch.where(torch.rand(K, M, generator=g) < 0.4,
                        torch.randint(0, M, (K, M), generator=g), torch.full((K, M), -1)).int().cuda()
    ws_bytes = lib.wcn_sort_workspace_bytes(M)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    r_tab = torch.empty(M, dtype=torch.int32, device="cuda")
    r_key = torch.empty(M, dtype=torch.int32, device="cuda")
    check(lib.wcn_sort_rows_by_table(table.data_ptr(), K, M, r_tab.data_ptr(), ws.data_ptr(), ws_bytes,
                                     _ops._stream()), "sort_rows_by_table")
    keys = _ops.mask_keys(table)
    check(lib.wcn_sort_rows_by_key(keys.data_ptr(), M, K, r_key.data_ptr(), ws.data_ptr(), ws_bytes,
                                   _ops._stream()), "sort_rows_by_key")
    torch.cuda.synchronize()
    assert torch.equal(r_tab, r_key)
    bits = (table >= 0).long()
    masks = (bits << torch.arange(K, device="cuda").view(K, 1)).sum(0)
    assert torch.equal(r_tab.long(), torch.argsort(_narrowed(masks, K), stable=True))


# ------------------------------------------------------------------------------------------------
# own all-reduce over NVLink peer memory (needs two GPUs: skipped on the one-GPU test box; run with
# `gpurun --gpus 2 -- python -m pytest tests -m gpu -k peer_allreduce`)
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_peer_allreduce_two_ranks_matches_nccl_and_replays_in_a_graph():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(root, "tools", "exp_peer_allreduce.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["graph_replay"] == "ok"
    for key in ("n442368", "n1000", "n4194308"):
        assert out[key]["exact_on_integers"] and out[key]["max_abs_vs_nccl"] < 1e-4


@pytest.mark.gpu
def test_peer_allreduce_rejects_bad_arguments():
    import ctypes
    from warpconvnet_b200._lib import lib
    arr = (ctypes.c_void_p * 2)(None, None)
    assert lib.wcn_peer_allreduce_flag_words() >= 128 * 16
    # world = 1 / n = 0: nothing to do
    assert lib.wcn_peer_allreduce_f32(arr, arr, 0, 1, 1024, ctypes.c_float(1.0), 8, None) == 0
    assert lib.wcn_peer_allreduce_f32(arr, arr, 0, 2, 0, ctypes.c_float(1.0), 8, None) == 0
    # null peer pointers, n not a multiple of 4, rank outside the world
    assert lib.wcn_peer_allreduce_f32(arr, arr, 0, 2, 1024, ctypes.c_float(1.0), 8, None) < 0
    assert lib.wcn_peer_allreduce_f32(arr, arr, 0, 2, 1023, ctypes.c_float(1.0), 8, None) < 0
    assert lib.wcn_peer_allreduce_f32(arr, arr, 2, 2, 1024, ctypes.c_float(1.0), 8, None) < 0


# ------------------------------------------------------------------------------------------------
# masked tile build, pinned arena
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_build_tiles_masked_equals_the_unmasked_build():
    """wcn_build_tiles_masked (rows read only the table entries their mask has) produces the same
    plan as wcn_build_tiles, for a submanifold and a strided map."""
    from warpconvnet_b200 import _ops
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    c = torch.from_numpy(surface_coords(150, 3))
    bc = torch.cat([torch.zeros(len(c), 1, dtype=torch.int32), c], 1).cuda()
    from warpconvnet_b200.geometry.coords.ops.stride import stride_coords
    out_bc, _ = stride_coords(bc, (2, 2, 2), n_batches=1)
    for in_c, out_c, stride, ks in ((bc, bc, (1, 1, 1), (3, 3, 3)), (bc, out_bc, (2, 2, 2), (2, 2, 2))):
        km = generate_kernel_map(in_c, out_c, stride, ks, build_plan=False)
        table = km.pair_table(out_c.shape[0])
        keys = _ops.mask_keys(table)
        masked = _ops.build_tile_plan(table, keys)          # passes the row masks
        plain = _ops.build_tile_plan(table, keys, key_bits=table.shape[0])  # same order, no masks
        assert torch.equal(masked.rows, plain.rows)
        assert torch.equal(masked.tile_nk, plain.tile_nk) and torch.equal(masked.tile_cum, plain.tile_cum)
        for t in range(masked.num_tiles):
            n = int(masked.tile_nk[t])
            assert torch.equal(masked.step_k[t, :n], plain.step_k[t, :n])
            assert torch.equal(masked.step_nbr[t, :n], plain.step_nbr[t, :n])


@pytest.mark.gpu
def test_deferred_read_back_survives_reuse_of_its_pinned_slot(monkeypatch):
    """The (offsets, status) read-back of a kernel map goes through a slot of one pinned arena; a
    map that is only resolved after its slot has been handed out again reads the device copy."""
    from warpconvnet_b200.geometry.coords.search import search_results as sr
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    monkeypatch.setattr(sr._PinnedArena, "SLOTS", 4)
    monkeypatch.setattr(sr, "_ARENA", None)
    maps, want = [], []
    for seed in range(10):
        c = torch.from_numpy(random_coords(400 + 37 * seed, 0.3, seed))
        bc = torch.cat([torch.zeros(len(c), 1, dtype=torch.int32), c], 1).cuda()
        maps.append(generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True))
        ref = okm.generate_kernel_map(bc.cpu().numpy(), bc.cpu().numpy(), (1, 1, 1), (3, 3, 3))
        want.append(np.asarray(ref["offsets"]))
    torch.cuda.synchronize()
    for km, w in zip(maps, want):          # the first six lost their slots to later maps
        assert np.array_equal(km.offsets.numpy().astype(np.int64), np.asarray(w).astype(np.int64))
    monkeypatch.setattr(sr, "_ARENA", None)


@pytest.mark.gpu
def test_peer_gradient_bucket_sections_two_ranks():
    """FlatGradBucket(peer=True, sections=k): sections reduced by post-accumulate hooks during
    backward equal the NCCL mean, eagerly and replayed from a CUDA graph (needs two GPUs)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29534",
           os.path.join(root, "tools", "exp_peer_bucket.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    for k in ("sections1", "sections3", "sections5"):
        assert out[k]["max_rel_err_vs_nccl_mean"] < 1e-5 and out[k]["timeouts"] == 0
