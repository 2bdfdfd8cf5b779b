# SPDX-License-Identifier: Apache-2.0
"""Bring-up script for the GPU box: runs each stage of the hot path in its own subprocess (a
trapped kernel poisons the CUDA context) and prints error patterns that help debug descriptor /
layout mistakes. Usage: python tools/gpu_check.py [stage ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STAGES = ["hash", "kmap", "dense_gemm", "fwd", "dgrad", "wgrad", "module", "shapes"]


def _mk(n=4000, seed=0):
    import numpy as np
    import torch
    from conftest import random_coords
    from oracle import kernel_map as okm
    c = random_coords(n, 0.3, seed)
    bc = okm.batch_indexed([c])
    return bc, torch.from_numpy(bc).cuda()


def stage_hash():
    import numpy as np
    import torch
    from oracle import kernel_map as okm
    from warpconvnet_b200.geometry.coords.search.packed_hashmap import PackedHashTable
    bc, bct = _mk()
    t = PackedHashTable.from_coords(bct)
    res = t.search(bct).cpu().numpy()
    assert (res == np.arange(len(bc))).all(), "self lookup failed"
    q = bct.clone(); q[:, 1] += 1000
    assert (t.search(q).cpu().numpy() == -1).all()
    print("hash ok")


def stage_kmap():
    import numpy as np
    import torch
    from oracle import kernel_map as okm
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    for n, stride in ((4000, 1), (100000, 1), (50000, 2)):
        bc, bct = _mk(n)
        if stride == 1:
            out_bc, out_t = bc, bct
        else:
            out_bc, _ = okm.stride_coords(bc, (stride,) * 3)
            out_t = torch.from_numpy(out_bc).cuda()
        t0 = time.time()
        km = generate_kernel_map(bct, out_t, (stride,) * 3, (3, 3, 3))
        torch.cuda.synchronize()
        ref = okm.generate_kernel_map(bc, out_bc, (stride,) * 3, (3, 3, 3))
        assert (km.offsets.numpy() == ref["offsets"]).all(), "offsets differ"
        assert (km._pair_table.cpu().numpy() == ref["pair_table"]).all(), "pair table differs"
        assert (km.in_maps.cpu().numpy() == ref["in_maps"]).all(), "in_maps differ"
        assert (km.out_maps.cpu().numpy() == ref["out_maps"]).all(), "out_maps differ"
        print(f"kmap ok n={n} stride={stride} L={int(km.offsets[-1])} ({time.time()-t0:.3f}s)")


def _report(name, got, ref):
    import torch
    got = got.double().cpu(); ref = ref.double().cpu()
    err = (got - ref).abs()
    rel = float(err.max() / ref.abs().max().clamp_min(1e-30))
    print(f"{name}: max_rel={rel:.3e} mean_abs_err={float(err.mean()):.3e} ref_absmean={float(ref.abs().mean()):.3e}")
    if rel > 5e-2:
        bad_rows = (err.max(dim=1).values > 0.05 * ref.abs().max()).nonzero().flatten()
        bad_cols = (err.max(dim=0).values > 0.05 * ref.abs().max()).nonzero().flatten()
        print(f"  bad rows: {bad_rows.numel()}/{got.shape[0]} first {bad_rows[:16].tolist()}")
        print(f"  bad cols: {bad_cols.numel()}/{got.shape[1]} first {bad_cols[:32].tolist()}")
        print("  got[0,:8]", got[0, :8].tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
        print("  nan count", int(torch.isnan(got).sum()))
    return rel


def stage_dense_gemm():
    """K=1 identity neighbour table: the kernel is a plain dense GEMM -> validates descriptors."""
    import torch
    from warpconvnet_b200 import _ops
    torch.manual_seed(0)
    for dtype in (torch.bfloat16, torch.float16, torch.float32):
        for (M, cin, cout) in ((128, 64, 128), (256, 128, 128), (300, 32, 64), (1000, 96, 256), (128, 16, 16)):
            x = torch.randn(M, cin, device="cuda").to(dtype)
            w = (torch.randn(1, cin, cout, device="cuda") / cin ** 0.5).to(dtype)
            table = torch.arange(M, dtype=torch.int32, device="cuda").view(1, M).contiguous()
            plan = _ops.build_tile_plan(table)
            img = _ops.weight_image(w.view(1, 1, cin, cout), 1, 1, cin, cout, False)
            y = _ops.gather_gemm(x, img, plan, 1, cin, cout)
            torch.cuda.synchronize()
            ref = x.double() @ w[0].double()
            _report(f"dense {dtype} M={M} {cin}->{cout}", y, ref)


def _conv_case(n, cin, cout, dtype, seed=0, stride=1):
    import numpy as np
    import torch
    from oracle import kernel_map as okm
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    bc, bct = _mk(n, seed)
    if stride == 1:
        out_t, out_bc = bct, bc
    else:
        out_bc, _ = okm.stride_coords(bc, (stride,) * 3)
        out_t = torch.from_numpy(out_bc).cuda()
    ks = (3, 3, 3) if stride == 1 else (2, 2, 2)
    km = generate_kernel_map(bct, out_t, (stride,) * 3, ks)
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    K = int(np.prod(ks))
    x = torch.randn(n, cin, generator=g).cuda().to(dtype)
    w = (torch.randn(K, cin, cout, generator=g) / (K * cin) ** 0.5).cuda().to(dtype)
    gy = torch.randn(len(out_bc), cout, generator=g).cuda().to(dtype)
    return km, x, w, gy, len(out_bc)


def stage_fwd():
    import torch
    from oracle import conv as oconv
    from warpconvnet_b200.nn.functional.sparse_conv import sparse_conv_forward
    for dtype in (torch.bfloat16, torch.float32):
        for (n, cin, cout, stride) in ((4000, 64, 128, 1), (20000, 128, 128, 1), (20000, 32, 64, 2)):
            km, x, w, gy, n_out = _conv_case(n, cin, cout, dtype, stride=stride)
            y = sparse_conv_forward(x, w, km, n_out)
            torch.cuda.synchronize()
            ref = oconv.forward(x.float().cpu(), w.float().cpu(), km.in_maps.cpu(), km.out_maps.cpu(), km.offsets, n_out)
            _report(f"fwd {dtype} n={n} {cin}->{cout} s={stride}", y, ref)


def stage_dgrad():
    import torch
    from oracle import conv as oconv
    from warpconvnet_b200.nn.functional.sparse_conv import sparse_conv_dgrad
    for dtype in (torch.bfloat16, torch.float32):
        for (n, cin, cout, stride) in ((4000, 64, 128, 1), (20000, 128, 128, 1), (20000, 32, 64, 2)):
            km, x, w, gy, n_out = _conv_case(n, cin, cout, dtype, stride=stride)
            dx = sparse_conv_dgrad(gy, w, km, n)
            torch.cuda.synchronize()
            rdx, _ = oconv.backward(gy.float().cpu(), x.float().cpu(), w.float().cpu(), km.in_maps.cpu(), km.out_maps.cpu(), km.offsets)
            _report(f"dgrad {dtype} n={n} {cin}->{cout} s={stride}", dx, rdx)


def stage_wgrad():
    import torch
    from oracle import conv as oconv
    from warpconvnet_b200.nn.functional.sparse_conv import sparse_conv_wgrad
    for dtype in (torch.bfloat16, torch.float32):
        for (n, cin, cout, stride) in ((4000, 64, 128, 1), (20000, 128, 128, 1), (20000, 32, 64, 2)):
            km, x, w, gy, n_out = _conv_case(n, cin, cout, dtype, stride=stride)
            dw = sparse_conv_wgrad(x, gy, tuple(w.shape), km)
            torch.cuda.synchronize()
            _, rdw = oconv.backward(gy.float().cpu(), x.float().cpu(), w.float().cpu(), km.in_maps.cpu(), km.out_maps.cpu(), km.offsets)
            _report(f"wgrad {dtype} n={n} {cin}->{cout} s={stride}", dw.reshape(-1, cout), rdw.reshape(-1, cout))


def stage_module():
    import torch
    from conftest import random_coords
    from oracle import conv as oconv
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules import SparseConv3d
    torch.manual_seed(0)
    coords = [torch.from_numpy(random_coords(3000, 0.3, s)) for s in (0, 1)]
    feats = [torch.randn(3000, 32) for _ in coords]
    v = Voxels(coords, feats, device="cuda")
    v.batched_features.batched_tensor.requires_grad_(True)
    conv1 = SparseConv3d(32, 64, 3).cuda()
    down = SparseConv3d(64, 64, 2, stride=2).cuda()
    up = SparseConv3d(64, 32, 2, stride=2, transposed=True).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        a = conv1(v)
        b = down(a)
        c = up(b, a)
    loss = c.feature_tensor.float().square().mean()
    loss.backward()
    torch.cuda.synchronize()
    print("module ok", a, b, c, float(loss), conv1.weight.grad.abs().mean().item(), up.weight.grad.abs().mean().item())


def stage_shapes():
    import torch
    from oracle import conv as oconv
    from warpconvnet_b200.nn.functional.sparse_conv import (sparse_conv_dgrad, sparse_conv_forward,
                                                            sparse_conv_wgrad)
    for (cin, cout) in ((4, 8), (16, 16), (48, 96), (256, 256), (192, 64), (64, 512)):
        km, x, w, gy, n_out = _conv_case(3000, cin, cout, torch.bfloat16)
        y = sparse_conv_forward(x, w, km, n_out)
        dx = sparse_conv_dgrad(gy, w, km, 3000)
        dw = sparse_conv_wgrad(x, gy, tuple(w.shape), km)
        torch.cuda.synchronize()
        args = (km.in_maps.cpu(), km.out_maps.cpu(), km.offsets)
        ref = oconv.forward(x.float().cpu(), w.float().cpu(), *args, n_out)
        rdx, rdw = oconv.backward(gy.float().cpu(), x.float().cpu(), w.float().cpu(), *args)
        _report(f"shape {cin}->{cout} fwd", y, ref)
        _report(f"shape {cin}->{cout} dgrad", dx, rdx)
        _report(f"shape {cin}->{cout} wgrad", dw.reshape(-1, cout), rdw.reshape(-1, cout))


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        globals()["stage_" + sys.argv[2]]()
        sys.exit(0)
    stages = sys.argv[1:] or STAGES
    failed = []
    for s in stages:
        print(f"===== stage {s} =====", flush=True)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", s], timeout=600)
        if r.returncode != 0:
            failed.append(s)
            print(f"stage {s} FAILED rc={r.returncode}", flush=True)
    print("FAILED:", failed)
    sys.exit(1 if failed else 0)
