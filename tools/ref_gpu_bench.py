# SPDX-License-Identifier: Apache-2.0
"""Same-box head-to-head against the UNMODIFIED reference's own GPU build (SURVEY.md §8d
``ref_auto_ms``). The reference (NVlabs/WarpConvNet) is built for sm_100a by
``baseline/build_ref.sh`` into ``baseline/_ref`` and driven here through its public API with its
default ``auto`` algorithm selection (its autotuner runs during the warm-up iterations, exactly
as in its own ``scripts/bench_unet_gb300.py:82-93``: 8 warm-up + 20 timed steps, CUDA events).

    python tools/ref_gpu_bench.py [c3s] [c3r] [kmap] [c4] [--scenes 8] [--out FILE]

Each section prints one JSON line {"section", "ref_ms", "ours_ms", ...}:
  c3s / c3r : SparseConv3d 3^3 128->128 bf16, ~200k voxels (surface / uniform random): fwd+bwd per
              step with the kernel map cached on the Voxels (the reference's protocol) and with a
              fresh Voxels every step (kernel map rebuilt)
  kmap      : the reference's generate_kernel_map (_C.cuhash) vs ours on the same coordinates:
              offsets equal, per-offset pair SETS equal (row order inside an offset is
              unspecified upstream)
  c4        : MinkUNet-14 shape, 8 scenes x ~300k voxels, AMP bf16, fwd+bwd+SGD
Nothing of this repo runs on the reference arm and vice versa; both arms see identical inputs.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
os.environ.setdefault("WARPCONVNET_BENCHMARK_CACHE_DIR", os.path.join(ROOT, "gpurun_out", "ref_cache"))
os.environ.setdefault("WARPCONVNET_AUTOTUNE_LOG", "false")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

PLANES = (32, 64, 128, 256, 128, 128, 96, 96)

# The reference's auto-tuner SEGFAULTS on this B200 inside _C.mask_gemm.dgrad when it reaches the
# last two candidates of its default dgrad pool (native "pcoff" dgrad tiles; the 13 candidates
# before them run: gpurun_out/r2d_ref_diag.log, profiles/r2_reference_gpu_head_to_head.md). To get
# a reference number at all, its dgrad pool is restricted — through its own public attribute
# SparseConv3d.dgrad_algo (a list of backend names, nn/modules/sparse_conv.py:130-137) — to every
# other family it would have tried, including the one that measured fastest in that sweep
# (mask_gemm_fwd_as_dgrad, 0.44 ms vs 0.48-0.49 ms for the excluded family's working tiles).
# Forward and wgrad keep the reference's default "auto".
REF_DGRAD_POOL = ["mask_gemm_fwd_as_dgrad", "cutlass_implicit_gemm", "cute_grouped",
                  "cutlass_grouped_hybrid", "explicit_gemm", "implicit_gemm"]


def pin_ref_dgrad_pool(module):
    import warpconvnet.nn.modules.sparse_conv as rsc
    n = 0
    for m in module.modules():
        if isinstance(m, rsc.SpatiallySparseConv):
            m.dgrad_algo = list(REF_DGRAD_POOL)
            n += 1
    return n


def surface(extent, seed):
    rng = np.random.RandomState(seed)
    a, b = rng.uniform(0, 2 * np.pi, size=2)
    u, v = np.meshgrid(np.arange(extent), np.arange(extent), indexing="ij")
    z = np.rint(12 * np.sin(2 * np.pi * u / 180 + a) + 8 * np.cos(2 * np.pi * v / 130 + b)) + 256
    return torch.from_numpy(np.stack([u.reshape(-1), v.reshape(-1), z.reshape(-1)], 1).astype(np.int32))


def random_cube(n, side, seed):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(side ** 3, generator=g)[:n]
    return torch.stack([idx // (side * side), (idx // side) % side, idx % side], 1).int()


def timed(fn, warmup=8, iters=20):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = (torch.cuda.Event(enable_timing=True) for _ in range(2))
    t0 = time.perf_counter()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters, (time.perf_counter() - t0) * 1e3 / iters


def load_ref():
    if not os.path.isdir(os.path.join(REF, "warpconvnet")):
        raise SystemExit(json.dumps({"unavailable": "baseline/_ref absent (run baseline/build_ref.sh)"}))
    sys.path.insert(0, REF)
    import warpconvnet  # noqa: F401
    from warpconvnet.geometry.types.voxels import Voxels as RVoxels
    from warpconvnet.nn.modules.sparse_conv import SparseConv3d as RConv
    return RVoxels, RConv


def conv_layer_section(name, coords, RVoxels, RConv):
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d
    dev = "cuda"
    n = len(coords)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, 128, generator=g).to(dev)
    w = (torch.randn(27, 128, 128, generator=g) * (27 * 128) ** -0.5).to(dev)
    gy = torch.randn(n, 128, generator=g).to(dev).bfloat16()
    c = coords.to(dev)
    out = {"section": name, "voxels": n}
    for arm, VoxT, ConvT in (("ref", RVoxels, RConv), ("ours", Voxels, SparseConv3d)):
        conv = ConvT(128, 128, 3, bias=False).to(dev)
        if arm == "ref":
            pin_ref_dgrad_pool(conv)
        with torch.no_grad():
            conv.weight.copy_(w)
        vox = VoxT([c], [x.bfloat16()])

        def step_cached():
            conv.weight.grad = None
            v = vox.replace(batched_features=vox.feature_tensor.detach().requires_grad_(True))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = conv(v)
            y.feature_tensor.backward(gy)

        def step_fresh():
            conv.weight.grad = None
            v = VoxT([c], [x.bfloat16().requires_grad_(True)])
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = conv(v)
            y.feature_tensor.backward(gy)

        def fwd_only():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                conv(vox)

        t0 = time.perf_counter()
        ms_c, wall_c = timed(step_cached)
        out[f"{arm}_warm_s"] = round(time.perf_counter() - t0, 1)
        ms_f, wall_f = timed(step_fresh, warmup=3)
        ms_fwd, _ = timed(fwd_only, warmup=3)
        out[f"{arm}_fwd_bwd_cached_map_ms"] = round(ms_c, 4)
        out[f"{arm}_fwd_bwd_fresh_map_ms"] = round(ms_f, 4)
        out[f"{arm}_fwd_only_cached_map_ms"] = round(ms_fwd, 4)
        out[f"{arm}_wall_fresh_ms"] = round(wall_f, 4)
        if arm == "ref":
            y_ref = None
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                y_ref = conv(vox).feature_tensor.float()
            c_ref = None
        else:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                y_our = conv(vox).feature_tensor.float()
    out["max_abs_diff_fwd_ref_vs_ours_over_max"] = float((y_ref - y_our).abs().max() / y_ref.abs().max())
    out["speedup_cached_map"] = round(out["ref_fwd_bwd_cached_map_ms"] / out["ours_fwd_bwd_cached_map_ms"], 2)
    out["speedup_fresh_map"] = round(out["ref_fwd_bwd_fresh_map_ms"] / out["ours_fwd_bwd_fresh_map_ms"], 2)
    return out


def kmap_section():
    """offsets / per-offset pair sets of the reference's _C.cuhash kernel map vs ours."""
    from warpconvnet.geometry.coords.search.torch_discrete import generate_kernel_map as ref_gkm
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map as our_gkm
    from warpconvnet_b200.geometry.coords.ops.stride import stride_coords
    res = {"section": "kmap", "cases": []}
    cases = [("c3s_k3_s1", surface(448, 0), 3, 1), ("c3r_k3_s1", random_cube(200000, 88, 0), 3, 1),
             ("surf_k2_s2", surface(317, 1), 2, 2), ("rand_k3_s2", random_cube(60000, 60, 2), 3, 2)]
    for name, c, ks, st in cases:
        bc = torch.cat([torch.zeros(len(c), 1, dtype=torch.int32), c], 1).cuda().contiguous()
        out_bc = bc if st == 1 else stride_coords(bc, (st,) * 3, n_batches=1)[0]
        rk = ref_gkm(bc, out_bc, (st,) * 3, (ks,) * 3)
        ok_ = our_gkm(bc, out_bc, (st,) * 3, (ks,) * 3)
        offs_equal = bool(torch.equal(rk.offsets.cpu().long(), ok_.offsets.cpu().long()))
        sets_equal = offs_equal
        if offs_equal:
            n_in = len(bc)
            for k in range(len(rk)):
                ri, ro = rk[k]
                oi, oo = ok_[k]
                a = torch.sort(ro.long() * n_in + ri.long()).values
                b = torch.sort(oo.long() * n_in + oi.long()).values
                if not torch.equal(a, b):
                    sets_equal = False
                    break
        res["cases"].append({"case": name, "voxels": len(c), "pairs": int(ok_.offsets[-1]),
                             "offsets_equal": offs_equal, "pair_sets_equal": sets_equal,
                             "identity_map_index_equal": rk.identity_map_index == ok_.identity_map_index})
    res["all_equal"] = all(x["offsets_equal"] and x["pair_sets_equal"] for x in res["cases"])
    return res


def c4_section(scenes, RVoxels):
    from minkunet14 import MinkUNet14
    from warpconvnet.models.mink_unet import MinkUNetBase
    from warpconvnet_b200.geometry.types.voxels import Voxels
    dev = "cuda"
    coords = [surface(548, s).to(dev) for s in range(scenes)]
    feats = [torch.randn(len(c), 3, device=dev) for c in coords]
    n = sum(len(c) for c in coords)
    out = {"section": "c4", "scenes": scenes, "voxels": n}
    for arm in ("ref", "ours"):
        torch.manual_seed(0)
        if arm == "ref":
            net = MinkUNetBase(3, 20, planes=PLANES, layers=(1,) * 8).to(dev)
            out["ref_convs_with_pinned_dgrad_pool"] = pin_ref_dgrad_pool(net)
            VoxT = RVoxels
        else:
            net = MinkUNet14(3, 20).to(dev)
            VoxT = Voxels
        out[f"{arm}_params_M"] = round(sum(p.numel() for p in net.parameters()) / 1e6, 3)
        opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)
        vox = VoxT(coords, feats)

        def step_cached():
            opt.zero_grad(set_to_none=True)
            v = vox.replace(batched_features=vox.feature_tensor.detach().requires_grad_(True))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = net(v)
            y.feature_tensor.float().pow(2).mean().backward()
            opt.step()

        def step_fresh():
            opt.zero_grad(set_to_none=True)
            v = VoxT(coords, feats)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = net(v)
            y.feature_tensor.float().pow(2).mean().backward()
            opt.step()

        torch.cuda.reset_peak_memory_stats()
        t0 = time.perf_counter()
        ms_c, wall_c = timed(step_cached, warmup=8, iters=20)
        out[f"{arm}_warm_plus_timed_s"] = round(time.perf_counter() - t0, 1)
        ms_f, wall_f = timed(step_fresh, warmup=3, iters=10)
        out[f"{arm}_step_cached_maps_ms"] = round(ms_c, 3)
        out[f"{arm}_step_cached_maps_wall_ms"] = round(wall_c, 3)
        out[f"{arm}_step_fresh_maps_ms"] = round(ms_f, 3)
        out[f"{arm}_step_fresh_maps_wall_ms"] = round(wall_f, 3)
        out[f"{arm}_peak_mem_GiB"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
        del net, opt, vox
        torch.cuda.empty_cache()
    out["speedup_cached_maps"] = round(out["ref_step_cached_maps_ms"] / out["ours_step_cached_maps_ms"], 2)
    out["speedup_fresh_maps"] = round(out["ref_step_fresh_maps_ms"] / out["ours_step_fresh_maps_ms"], 2)
    out["voxels_per_s_ref"] = n / (out["ref_step_cached_maps_ms"] * 1e-3)
    out["voxels_per_s_ours"] = n / (out["ours_step_cached_maps_ms"] * 1e-3)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sections", nargs="*", default=["kmap", "c3s", "c3r", "c4"])
    ap.add_argument("--scenes", type=int, default=8)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_gpu.jsonl"))
    args = ap.parse_args()
    RVoxels, RConv = load_ref()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    lines = []
    for sec in args.sections:
        t0 = time.perf_counter()
        try:
            if sec == "kmap":
                r = kmap_section()
            elif sec == "c3s":
                r = conv_layer_section("c3s", surface(448, 0), RVoxels, RConv)
            elif sec == "c3r":
                r = conv_layer_section("c3r", random_cube(200000, 88, 0), RVoxels, RConv)
            elif sec == "c4":
                r = c4_section(args.scenes, RVoxels)
            else:
                raise ValueError(sec)
        except Exception as exc:  # keep going: one failing section must not hide the others
            import traceback
            r = {"section": sec, "error": f"{type(exc).__name__}: {exc}",
                 "trace": traceback.format_exc()[-1500:]}
        r["section_wall_s"] = round(time.perf_counter() - t0, 1)
        r["gpu"] = torch.cuda.get_device_name(0)
        print(json.dumps(r), flush=True)
        lines.append(r)
        with open(args.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
