#!/usr/bin/env python
# SPDX-License-Identifier: Apache-2.0
"""bench.py — SparseConv3d fwd+bwd voxels/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dist S|R] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE config C3 — one SparseConv3d 3^3, 128 -> 128 channels, bf16,
~200k active voxels per GPU (surface-like distribution S: 448x448 height field = 200 704 voxels;
--dist R: 200 000 uniformly random voxels at 30 % occupancy). One step = the whole hot path over one
batch: kernel-map build (hash + 27-offset probe + CSR) + tile plan + forward AB_gather_scatter +
dgrad ABt_gather_scatter + wgrad AtB_gather_gather (+ one NCCL all-reduce of dW when N > 1).

  value : voxels/s with the inputs already resident in HBM
  e2e   : the same through the public API (Voxels -> SparseConv3d -> backward) from pinned HOST
          buffers, H2D and D2H inside the timed region
  roofline     : forward gather-GEMM kernel alone (CUDA events around its launch in the timed
                 steps): algorithmic FLOPs 2*L*Cin*Cout / time vs the measured bf16 peak
  cpu_baseline : the oracle port of the reference's explicit gather-matmul-scatter
                 (oracle/conv.py, fp32, all host threads) on the same workload, rank 0, N = 1 only

--impl reference runs ONLY that CPU port (the reference is a Python/CUDA library whose CUDA
extension cannot be built here and /root/reference does not exist on the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CIN, COUT, KS = 128, 128, 3
K = KS ** 3


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
def make_coords(dist: str, seed: int) -> np.ndarray:
    if dist == "S":
        rng = np.random.RandomState(seed)
        a, b = rng.uniform(0, 2 * np.pi, size=2)
        u, v = np.meshgrid(np.arange(448), np.arange(448), indexing="ij")
        z = np.rint(12 * np.sin(2 * np.pi * u / 180 + a) + 8 * np.cos(2 * np.pi * v / 130 + b)) + 256
        return np.stack([u.reshape(-1), v.reshape(-1), z.reshape(-1)], 1).astype(np.int32)
    n, side = 200000, 88
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(side ** 3, generator=g)[:n].numpy()
    return np.stack([idx // (side * side), (idx // side) % side, idx % side], 1).astype(np.int32)


def make_tensors(n: int, seed: int):
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, CIN, generator=g)
    w = torch.randn(K, CIN, COUT, generator=g) * (K * CIN) ** -0.5
    gy = torch.randn(n, COUT, generator=g)
    return x, w, gy


def workload_name(dist: str, n: int, world: int) -> str:
    shape = "surface height field 448^2" if dist == "S" else "uniform random, 30% occupancy"
    return (f"C3-{dist}: SparseConv3d 3^3 {CIN}->{COUT} bf16, {n} voxels/GPU ({shape}), "
            "step = kernel-map build + tile plan + fwd AB + dgrad ABt + wgrad AtB"
            + (" + NCCL all-reduce(dW)" if world > 1 else ""))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tflops": float(p["bf16_tflops_sustained"]), "hbm": float(p["hbm_gbs"]),
                "burst": float(p.get("bf16_tflops", 0)) or float(p["bf16_tflops_sustained"]),
                "which": "measured (MEASURED_PEAKS.json)"}
    return {"tflops": 1400.0, "hbm": 6650.0, "burst": 1640.0,
            "which": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# CPU port of the reference's explicit path (the oracle) — baseline legs only
# ------------------------------------------------------------------------------------------------
def cpu_port_step(coords: np.ndarray, x, w, gy):
    """One pass of the hot path on the host: kernel map (oracle/kernel_map.py) + explicit
    gather-matmul-scatter fwd + bwd in fp32 (oracle/conv.py). Returns (seconds, L)."""
    from oracle import conv as oconv
    from oracle import kernel_map as okm
    t0 = time.perf_counter()
    bc = okm.batch_indexed([coords])
    km = okm.generate_kernel_map(bc, bc, (1, 1, 1), (KS,) * 3)
    args = (km["in_maps"], km["out_maps"], km["offsets"])
    oconv.forward(x, w, *args, len(bc), dtype=torch.float32)
    oconv.backward(gy, x, w, *args, dtype=torch.float32)
    return time.perf_counter() - t0, int(km["offsets"][-1])


_ALL_CPUS = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None


def run_cpu_baseline(dist: str, reps: int, warmup: int, budget_s: float = 40.0):
    if _ALL_CPUS is not None:
        os.sched_setaffinity(0, _ALL_CPUS)   # the CPU arm uses every host core (undo the NUMA binding)
    cores = len(_ALL_CPUS) if _ALL_CPUS is not None else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    coords = make_coords(dist, 0)
    x, w, gy = make_tensors(len(coords), 0)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + reps):
        t, _ = cpu_port_step(coords, x, w, gy)
        if i >= warmup:
            times.append(t)
        if time.perf_counter() - t_begin > budget_s and times:
            break
    t = float(np.mean(times))
    return {"value": len(coords) / t, "unit": "voxels/s", "cores": cores, "kind": "port",
            "sample": f"full C3-{dist} workload ({len(coords)} voxels, fp32, kernel map + fwd + "
                      f"dgrad + wgrad), mean of {len(times)} passes after {warmup} warm-up",
            "seconds_per_pass": t}, len(coords)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / power / throttle reasons of one GPU through NVML every 5 ms in a thread
    (the timed region is tens of milliseconds, too short for an `nvidia-smi -lms` process)."""

    def __init__(self, index: int):
        import threading
        self.rows, self.stop_flag, self.err = [], False, None
        self.period = 0.0005  # seconds between NVML samples (raised for the host-paced e2e loop)
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  self.sm_max,
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                break
            time.sleep(self.period)

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        self.stop_flag = True
        self.t.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(n for n, bit in names.items() if any(r[3] & bit for r in self.rows))
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(r[1] for r in self.rows)) if sm else None,
                "samples": len(sm),
                "power_w_max": float(max(r[2] for r in self.rows)) if sm else None,
                "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# BASELINE config 4: MinkUNet-14 shape, 8 scenes x ~300k voxels, strong scaling over the ranks
# ------------------------------------------------------------------------------------------------
C4_SCENES, C4_EXTENT = 8, 548


def run_c4(dev, rank, world, dist, steps=10, warmup=3):
    """MinkUNet-14 shape (tools/minkunet14.py = the reference's MinkUNetBase(3, 20,
    planes=(32,64,128,256,128,128,96,96), layers=(1,)*8), models/mink_unet.py:259-336), global batch
    of 8 surface scenes (300 304 voxels each) sharded round-robin over the ranks (STRONG scaling:
    8 / 4 / 2 / 1 scenes per rank), AMP bf16, fwd + bwd + one flat all-reduce of the fp32 gradients
    + SGD(momentum) step. Kernel maps are rebuilt from the coordinates in EVERY step (fresh Voxels),
    which the reference's own protocol (scripts/bench_unet_gb300.py:82-93) does not do — its maps
    stay cached on the Voxels across iterations; `maps_cached_ms` times that protocol too.
    Timed three ways: eager launches (CUDA events + wall clock), and the whole step replayed as ONE
    CUDA graph (sizes from the SizeTape of the eager warm-up; valid because the geometry is fixed)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from minkunet14 import MinkUNet14, surface_scene
    from warpconvnet_b200.dist import FlatGradBucket, shard_scenes
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.utils.graph import capture_step

    mine = shard_scenes(C4_SCENES, rank, world)
    coords = [surface_scene(C4_EXTENT, s).to(dev) for s in mine]
    g = torch.Generator().manual_seed(100 + rank)
    feats = [torch.randn(len(c), 3, generator=g).to(dev) for c in coords]
    n_local = sum(len(c) for c in coords)
    torch.manual_seed(0)                                   # identical initial weights on all ranks
    net = MinkUNet14(3, 20).to(dev)
    opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)
    bucket = None
    if world > 1:
        try:  # gradients in NVLink peer-mapped memory, reduced by the library's own kernel
            bucket = FlatGradBucket(net.parameters(), peer=True, sections=4)
        except Exception:  # symmetric memory unavailable: NCCL
            bucket = FlatGradBucket(net.parameters())
    cached_vox = Voxels(coords, feats)
    # replace() copies the attribute dict, so a kernel-map cache only survives across steps when it
    # exists on the template object BEFORE the first replace (otherwise every step's copy creates
    # its own and the next step starts empty again — which is also what happens in the reference's
    # scripts/bench_unet_gb300.py, whose "cached" maps are in fact rebuilt every iteration)
    from warpconvnet_b200.geometry.coords.search.cache import IntSearchCache
    cached_vox._extra_attributes["_cache"] = IntSearchCache()

    def make_step(fresh_maps: bool):
        def step():
            x = Voxels(coords, feats) if fresh_maps else cached_vox.replace(
                batched_features=cached_vox.feature_tensor.detach())
            if world > 1:
                bucket.zero()          # .grad stay views into the flat all-reduce buffer
            else:
                opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(x)
            loss = out.feature_tensor.float().square().mean()
            loss.backward()
            if world > 1:
                bucket.all_reduce(average=True)
            opt.step()
            return loss
        return step

    step = make_step(True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k, w):
        for _ in range(w):
            fn()
        sync_all()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        for _ in range(k):
            fn()
        e.record()
        sync_all()
        wall = (time.perf_counter() - t0) * 1e3 / k
        t = torch.tensor([s.elapsed_time(e) / k, wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    eager_ms, eager_wall = timed(step, steps, warmup)
    cached_ms, _ = timed(make_step(False), steps, 2)
    graph_ms = graph_note = c4_clocks = None
    try:
        graph, tape, _ = capture_step(step, warmup=1)
        c4_sampler = ClockSampler(dev.index if dev.index is not None else 0) if rank == 0 else None
        graph_ms, graph_wall = timed(graph.replay, max(steps, 30), 2)   # >= 1 s of replays
        c4_clocks = c4_sampler.stop() if c4_sampler is not None else None
        tape.verify()
    except Exception as exc:  # pragma: no cover
        graph_note = f"graph capture failed: {type(exc).__name__}: {str(exc)[:200]}"
        torch.cuda.synchronize()
    counts = torch.zeros(world, device=dev, dtype=torch.float64)
    counts[rank] = n_local
    if world > 1:
        dist.all_reduce(counts)
    total = float(counts.sum().item())
    best = graph_ms if graph_ms is not None else eager_ms
    out = {
        "workload": ("C4: MinkUNet-14 shape (%.2f M params), global batch %d scenes x 300 304 voxels "
                     "(surface height field 548^2), AMP bf16, fwd+bwd+all-reduce(grads)+SGD, "
                     "kernel maps rebuilt every step" % (sum(p.numel() for p in net.parameters()) / 1e6,
                                                          C4_SCENES)),
        "scaling": "strong", "n_gpus": world, "scenes_per_rank": len(mine),
        "voxels_per_rank": [int(v) for v in counts.tolist()], "voxels_total": int(total),
        "ms_per_step": best, "value": total / (best * 1e-3), "unit": "voxels/s",
        "graph_replay_ms": graph_ms, "eager_ms": eager_ms, "eager_wall_ms": eager_wall,
        "maps_cached_eager_ms": cached_ms,
        "timing": "CUDA events around K steps, max over ranks; graph_replay = whole step "
                  "(coordinate hierarchy, 9 kernel maps, 24 convs fwd+bwd, norms, all-reduce, SGD) "
                  "as one CUDA graph; eager_wall = host wall clock of the same eager steps",
        "grad_allreduce_bytes": int(bucket.flat.numel() * 4) if bucket is not None else 0,
        "grad_allreduce": ("none" if bucket is None else
                           "own peer-memory kernel (wcn_peer_allreduce_f32), 4 sections launched by "
                           "post-accumulate hooks during backward" if bucket.peer is not None
                           else "NCCL"),
        "peak_mem_GiB": torch.cuda.max_memory_allocated() / 2 ** 30,
        "steps": steps, "warmup": warmup,
        "clocks_during_graph_replays": c4_clocks,
    }
    if graph_note:
        out["graph_note"] = graph_note
    del net, opt, bucket, cached_vox
    torch.cuda.empty_cache()
    return out


def run_ref_gpu(sections, timeout_s):
    """Same-box head-to-head against the UNMODIFIED reference's own GPU build (baseline/_ref, built
    by baseline/build_ref.sh) through tools/ref_gpu_bench.py, in a SUBPROCESS: the reference's
    auto-tuner is known to segfault on B200 (tools/ref_gpu_bench.py: REF_DGRAD_POOL) and must not
    be able to take this process down. Returns the tool's JSON lines keyed by section."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref", "warpconvnet")
    if not os.path.isdir(ref_dir):
        return {"unavailable": "baseline/_ref absent (bash baseline/build_ref.sh builds it, ~15 min)"}
    out_path = os.path.join(tempfile.gettempdir(), f"wcn_ref_gpu_{os.getpid()}.jsonl")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "ref_gpu_bench.py"), *sections, "--out", out_path]
    t0 = time.perf_counter()
    try:
        proc = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
        note = None if proc.returncode == 0 else f"exit code {proc.returncode}"
    except subprocess.TimeoutExpired:
        note = f"timed out after {timeout_s} s"
    res = {"tool": "tools/ref_gpu_bench.py " + " ".join(sections),
           "protocol": "reference's own scripts/bench_unet_gb300.py protocol: 8 warm-up steps (its "
                       "auto-tuner runs there) + 20 timed, CUDA events; both arms in one process on "
                       "identical inputs; *_ms are per step",
           "wall_s": round(time.perf_counter() - t0, 1)}
    if note:
        res["note"] = note
    if os.path.exists(out_path):
        for line in open(out_path):
            try:
                r = json.loads(line)
                res[r.pop("section")] = r
            except Exception:
                pass
        os.remove(out_path)
    return res


def run_side_blocks(dev, flush):
    """C3-R, C2 and C5 on one GPU (SURVEY.md §8d says R must be reported next to S)."""
    from warpconvnet_b200 import _ops
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    from warpconvnet_b200.nn.functional.sparse_conv import (sparse_conv_dgrad, sparse_conv_forward,
                                                            sparse_conv_wgrad)

    def timed(fn, k=10, w=3):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(k):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return float(np.mean([a.elapsed_time(b) for a, b in evs]))

    peaks = load_peaks()
    blocks = {}
    # ---- C3-R: uniform-random occupancy (adversarial for output-stationary designs) ----------
    c = make_coords("R", 0)
    n = len(c)
    bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), c], 1)).to(dev)
    x_h, w_h, gy_h = make_tensors(n, 0)
    x, w, gy = x_h.to(dev).bfloat16(), w_h.to(dev).bfloat16(), gy_h.to(dev).bfloat16()
    km = generate_kernel_map(bc, bc, (1, 1, 1), (KS,) * 3, same_coords=True)
    L = int(km.offsets[-1])
    plan = km.fwd_plan(n)
    img, img_t = _ops.weight_image_pair(w.view(K, 1, CIN, COUT), K, 1, CIN, COUT, w.dtype)
    t_map = timed(lambda: generate_kernel_map(bc, bc, (1, 1, 1), (KS,) * 3, same_coords=True))
    t_f = timed(lambda: _ops.gather_gemm(x, img, plan, 1, CIN, COUT))
    bplan, kflip = km.bwd_plan(n)
    t_d = timed(lambda: _ops.gather_gemm(gy, img_t, bplan, 1, COUT, CIN, kflip=kflip))
    t_w = timed(lambda: sparse_conv_wgrad(x, gy, (K, CIN, COUT), km))
    fl = 2.0 * L * CIN * COUT
    blocks["c3_R"] = {
        "workload": workload_name("R", n, 1), "pairs_L": L,
        "kernel_map_plus_plan_ms": t_map, "fwd_ms": t_f, "dgrad_ms": t_d, "wgrad_ms": t_w,
        "step_ms_sum": t_map + t_f + t_d + t_w, "voxels_per_s": n / ((t_map + t_f + t_d + t_w) * 1e-3),
        "fwd_TFLOPs": fl / (t_f * 1e-3) / 1e12, "fwd_frac_of_sustained_peak": fl / (t_f * 1e-3) / 1e12 / peaks["tflops"],
        "plan_steps": int(plan.tile_nk.sum().item()), "tile_rows": plan.tile_rows}
    # ---- C2: 64 -> 128, ~100k voxels, forward only ------------------------------------------
    rng_c = make_coords("S", 0)
    sel = (rng_c[:, 0] < 317) & (rng_c[:, 1] < 317)
    c2 = rng_c[sel]
    n2 = len(c2)
    bc2 = torch.from_numpy(np.concatenate([np.zeros((n2, 1), np.int32), c2], 1)).to(dev)
    g = torch.Generator().manual_seed(2)
    x2 = torch.randn(n2, 64, generator=g).to(dev).bfloat16()
    w2 = (torch.randn(27, 64, 128, generator=g) * (27 * 64) ** -0.5).to(dev).bfloat16()
    km2 = generate_kernel_map(bc2, bc2, (1, 1, 1), (3, 3, 3), same_coords=True)
    L2 = int(km2.offsets[-1])
    t_map2 = timed(lambda: generate_kernel_map(bc2, bc2, (1, 1, 1), (3, 3, 3), same_coords=True))
    t_f2 = timed(lambda: sparse_conv_forward(x2, w2, km2, n2))
    blocks["c2"] = {"workload": "C2: SparseConv3d 3^3 64->128 bf16, %d voxels (S), fwd only" % n2,
                    "pairs_L": L2, "kernel_map_plus_plan_ms": t_map2, "fwd_ms": t_f2,
                    "voxels_per_s": n2 / ((t_map2 + t_f2) * 1e-3),
                    "fwd_TFLOPs": 2.0 * L2 * 64 * 128 / (t_f2 * 1e-3) / 1e12}
    # ---- C5 (one rank's share of the 8-GPU config): group conv 512 -> 512, groups = 64, and
    # PointConv(64, 64, knn_k = 16) on 125k points, bf16 --------------------------------------
    from warpconvnet_b200.geometry.coords.search.search_configs import RealSearchConfig
    from warpconvnet_b200.geometry.types.points import Points
    from warpconvnet_b200.nn.modules.point_conv import PointConv
    sel5 = (rng_c[:, 0] < 354) & (rng_c[:, 1] < 354)
    c5 = rng_c[sel5]
    n5 = len(c5)
    bc5 = torch.from_numpy(np.concatenate([np.zeros((n5, 1), np.int32), c5], 1)).to(dev)
    km5 = generate_kernel_map(bc5, bc5, (1, 1, 1), (3, 3, 3), same_coords=True)
    x5 = torch.randn(n5, 512, device=dev).bfloat16()
    gy5 = torch.randn(n5, 512, device=dev).bfloat16()
    w5 = (torch.randn(27, 64, 8, 8, device=dev) * (27 * 8) ** -0.5).bfloat16()

    def gstep():
        sparse_conv_forward(x5, w5, km5, n5, groups=64)
        sparse_conv_dgrad(gy5, w5, km5, n5, groups=64)
        sparse_conv_wgrad(x5, gy5, tuple(w5.shape), km5, groups=64)

    t_g = timed(gstep)
    npts = 125000
    gp = torch.Generator().manual_seed(5)
    pts = torch.rand(npts, 3, generator=gp).to(dev)
    pf = torch.randn(npts, 64, generator=gp).to(dev)
    offs = torch.tensor([0, npts], dtype=torch.int64)
    pconv = PointConv(64, 64, RealSearchConfig("knn", knn_k=16)).to(dev)
    t_knn = timed(lambda: _ops.knn_search(pts, offs, pts, offs, 16))

    def pstep():
        pc = Points(pts, pf.detach().requires_grad_(True), offsets=offs)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out_pc = pconv(pc)
        out_pc.feature_tensor.float().square().mean().backward()

    t_pc = timed(pstep, k=5, w=2)
    blocks["c5_rank_share"] = {
        "workload": "C5 (1/8 of the 8-GPU config): SparseConv3d(512,512,3,groups=64) on %d voxels "
                    "fwd+dgrad+wgrad; PointConv(64,64,knn_k=16) on %d points fwd+bwd, bf16" % (n5, npts),
        "group_conv_fwd_bwd_ms": t_g, "group_conv_voxels_per_s": n5 / (t_g * 1e-3),
        "knn_ms": t_knn, "pointconv_fwd_bwd_ms": t_pc, "pointconv_points_per_s": npts / (t_pc * 1e-3)}
    return blocks


def run_c5(dev, rank, world, dist, flush):
    """BASELINE config C5 at its stated scale, STRONG scaling over the ranks: 8 scenes x 125 316
    voxels (1.0 M voxels) through SparseConv3d(512, 512, 3, groups=64) fwd + dgrad + wgrad (+ the
    all-reduce of dW), and 8 clouds x 125 000 points (1.0 M points) through
    PointConv(64, 64, knn_k=16) fwd + bwd; every rank owns 8 / world scenes and clouds."""
    from warpconvnet_b200.dist import shard_scenes
    from warpconvnet_b200.geometry.coords.search.search_configs import RealSearchConfig
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    from warpconvnet_b200.geometry.types.points import Points
    from warpconvnet_b200.nn.functional.sparse_conv import (sparse_conv_dgrad, sparse_conv_forward,
                                                            sparse_conv_wgrad)
    from warpconvnet_b200.nn.modules.point_conv import PointConv
    mine = shard_scenes(8, rank, world)

    def timed(fn, k=5, w=2):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        evs = []
        for _ in range(k):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        t = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in evs]))], device=dev,
                         dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    base = make_coords("S", 0)
    c5 = base[(base[:, 0] < 354) & (base[:, 1] < 354)]
    per = len(c5)
    bc = torch.from_numpy(np.concatenate(
        [np.concatenate([np.full((per, 1), b, np.int32), c5], 1) for b in range(len(mine))], 0)).to(dev)
    n_local = bc.shape[0]
    km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
    g = torch.Generator(device=dev).manual_seed(50 + rank)
    x = torch.randn(n_local, 512, device=dev, generator=g).bfloat16()
    gy = torch.randn(n_local, 512, device=dev, generator=g).bfloat16()
    w = (torch.randn(27, 64, 8, 8, device=dev, generator=g) * (27 * 8) ** -0.5).bfloat16()

    def gstep():
        sparse_conv_forward(x, w, km, n_local, groups=64)
        sparse_conv_dgrad(gy, w, km, n_local, groups=64)
        dw = sparse_conv_wgrad(x, gy, tuple(w.shape), km, groups=64)
        if world > 1:
            dist.all_reduce(dw)

    t_g = timed(gstep)
    npts = 125000
    gp = torch.Generator().manual_seed(5 + rank)
    pts = torch.rand(npts * len(mine), 3, generator=gp).to(dev)
    pf = torch.randn(npts * len(mine), 64, generator=gp).to(dev)
    offs = torch.arange(len(mine) + 1, dtype=torch.int64) * npts
    torch.manual_seed(0)
    pconv = PointConv(64, 64, RealSearchConfig("knn", knn_k=16)).to(dev)

    def pstep():
        pconv.zero_grad(set_to_none=True)
        pc = Points(pts, pf.detach().requires_grad_(True), offsets=offs)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out_pc = pconv(pc)
        out_pc.feature_tensor.float().square().mean().backward()
        if world > 1:
            for q in pconv.parameters():
                if q.grad is not None:
                    dist.all_reduce(q.grad)

    t_pc = timed(pstep, k=3, w=2)
    out = {"workload": "C5: SparseConv3d(512,512,3,groups=64) on 8 scenes x %d voxels fwd+dgrad+wgrad"
                       "(+all-reduce dW); PointConv(64,64,knn_k=16) on 8 clouds x %d points fwd+bwd"
                       "(+all-reduce grads), bf16, kernel map / kNN built once for the group conv, "
                       "kNN rebuilt every PointConv step" % (per, npts),
           "scaling": "strong", "n_gpus": world, "scenes_per_rank": len(mine),
           "voxels_total": 8 * per, "points_total": 8 * npts,
           "group_conv_fwd_bwd_ms": t_g, "group_conv_voxels_per_s": 8 * per / (t_g * 1e-3),
           "pointconv_fwd_bwd_ms": t_pc, "pointconv_points_per_s": 8 * npts / (t_pc * 1e-3),
           "timing": "CUDA events per step, mean over steps, max over ranks; L2 flushed before each step"}
    del x, gy, km, pts, pf, pconv
    torch.cuda.empty_cache()
    return out


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process (and therefore the pinned host buffers it allocates afterwards) to the CPU
    cores of the NUMA node its GPU hangs off, so the per-step H2D copies of the e2e loop do not
    cross the socket interconnect. Best effort: returns a description or the reason it was skipped."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = bus.lower()[-12:]                       # 0000:xx:yy.z
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read().strip())
        if node < 0:
            return {"numa_node": None, "note": "platform reports no NUMA affinity for the GPU"}
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return {"numa_node": node, "note": "no allowed CPU on that node"}
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception as exc:  # pragma: no cover - depends on the box
        return {"numa_node": None, "note": f"{type(exc).__name__}: {str(exc)[:80]}"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from warpconvnet_b200 import _ops
    from warpconvnet_b200._lib import lib
    from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
    from warpconvnet_b200.geometry.types.voxels import Voxels
    from warpconvnet_b200.nn.functional.sparse_conv import sparse_conv_dgrad, sparse_conv_wgrad
    from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d

    coords = make_coords(args.dist, seed=rank)  # every rank owns its own scene (weak scaling)
    n = len(coords)
    x_h, w_h, gy_h = make_tensors(n, seed=rank)
    bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), coords], 1)).to(dev)
    x = x_h.to(dev).bfloat16()
    w = w_h.to(dev).bfloat16()
    gy = gy_h.to(dev).bfloat16()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The only collective of the path: the sum of dW over ranks. Default: this library's own
    # peer-memory all-reduce kernel (csrc/peer_allreduce.cu) on a side stream, reducing the
    # symmetric buffer wgrad wrote into, under the dgrad kernel; --collective nccl keeps the
    # library collective for comparison.
    par = ar_stream = None
    collective = "none"
    if world > 1:
        collective = args.collective
        if collective == "peer":
            why = ""
            try:
                from warpconvnet_b200.dist import PeerAllReduce
                par = PeerAllReduce(K * CIN * COUT, dev)
                ar_stream = torch.cuda.Stream(device=dev)
            except Exception as exc:  # symmetric memory not available on this box
                why = f"{type(exc).__name__}: {str(exc)[:120]}"
                par = None
            ok = torch.tensor([0 if par is None else 1], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # every rank takes the same path
            if int(ok.item()) == 0:
                par = None
                collective = f"nccl (peer memory unavailable on some rank{': ' + why if why else ''})"

    def step():
        """whole hot path, inputs resident in HBM; nothing in it synchronises with the host"""
        km = generate_kernel_map(bc, bc, (1, 1, 1), (KS,) * 3, same_coords=True)
        plan = km.fwd_plan(n)
        # forward + dgrad weight images in one launch (what SparseConv3d's autograd function does)
        img, img_t = _ops.weight_image_pair(w.view(K, 1, CIN, COUT), K, 1, CIN, COUT, w.dtype)
        y = _ops.gather_gemm(x, img, plan, 1, CIN, COUT)           # forward AB_gather_scatter
        # wgrad AtB_gather_gather (into the peer-mapped buffer when the own collective is used)
        dw = sparse_conv_wgrad(x, gy, (K, CIN, COUT), km,
                               out=None if par is None else par.buffer)
        # all-reduce of dW under dgrad (what DDP does with the rest of backward). Own kernel: dgrad
        # is enqueued FIRST, then the all-reduce on a side stream that forked after wgrad — its
        # small CTAs (128 threads, no shared memory) slot in next to the resident dgrad CTAs; the
        # other order makes the dgrad CTAs queue behind them (profiles/r2t_peer_allreduce.md)
        work = None
        bplan, kflip = km.bwd_plan(n)                              # submanifold: fwd plan, k flipped
        if par is not None:
            cur = torch.cuda.current_stream()
            ar_stream.wait_stream(cur)
        elif world > 1:
            work = dist.all_reduce(dw, async_op=True)
        dx = _ops.gather_gemm(gy, img_t, bplan, 1, COUT, CIN, kflip=kflip)  # dgrad ABt_gather_scatter
        if par is not None:
            with torch.cuda.stream(ar_stream):
                par.all_reduce_()
            cur.wait_stream(ar_stream)
        if work is not None:
            work.wait()
        return km, plan, img, y, dx, dw

    def timed(fn, steps, warmup):
        """W warm-up calls, then K timed calls: L2 flush (outside the events), start event, fn,
        stop event; barrier + synchronize on both sides; mean over steps, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        l0 = lib.wcn_launch_count()
        wall0 = time.perf_counter()
        for _ in range(steps):
            flush.fill_(1)  # L2 flush (256 MiB write) before every timed step
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        barrier()
        wall = time.perf_counter() - wall0
        launches = lib.wcn_launch_count() - l0
        ms = sum(s.elapsed_time(e) for s, e in evs) / steps
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, wall

    # ---- device-resident number: the step is captured ONCE into a CUDA graph (no host sync in
    # the path, so the whole map build + plan + fwd + dgrad + wgrad is capturable) and replayed;
    # eager launches of the same step are reported next to it ------------------------------------
    for _ in range(3):
        km, plan, img, y, dx, dw = step()
    torch.cuda.synchronize()
    L = int(km.offsets[-1])
    launches_per_step = None
    graph = None
    graph_note = None
    if not args.no_graph:
        try:
            l0 = lib.wcn_launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="relaxed"):
                g_out = step()  # includes the NCCL all-reduce of dW when world > 1
            launches_per_step = lib.wcn_launch_count() - l0
            graph.replay()
            torch.cuda.synchronize()
        except Exception as exc:  # pragma: no cover - depends on the NCCL / driver combination
            graph = None
            graph_note = f"graph capture failed ({type(exc).__name__}), eager launches timed"
            torch.cuda.synchronize()
    # NVML clock / power / throttle-reason samples are taken from here to the end of the e2e loop:
    # the headline region alone is ~10 ms, shorter than a handful of NVML round trips
    sampler = ClockSampler(local) if rank == 0 else None
    if graph is not None:
        ms, _, wall = timed(graph.replay, args.steps, args.warmup)
        launches = launches_per_step * args.steps
        eager_ms, _, _ = timed(step, args.steps, args.warmup)
    else:
        ms, launches, wall = timed(step, args.steps, args.warmup)
        eager_ms = ms
    total_vox = torch.tensor([n], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_vox)
    total_vox = float(total_vox.item())

    # ---- the dominant kernel alone (forward gather-GEMM): same plan / inputs, L2 flushed before
    # every launch, events on the launching stream; the flush keeps the GPU busy while the host
    # enqueues, so the events bracket the kernel and not the launch latency ----------------------
    gemm_ms, _, _ = timed(lambda: _ops.gather_gemm(x, img, plan, 1, CIN, COUT, out=y),
                          args.steps, args.warmup)

    # ---- per-phase breakdown (untimed for the headline; same step, events between phases) -----
    def breakdown(reps=5):
        names = ["kernel_map", "tile_plan", "weight_image", "fwd_gemm", "dgrad", "wgrad"]
        acc = np.zeros(len(names))
        for _ in range(reps):
            flush.fill_(1)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            ev[0].record()
            km_ = generate_kernel_map(bc, bc, (1, 1, 1), (KS,) * 3, same_coords=True)
            ev[1].record()
            plan_ = km_.fwd_plan(n)
            ev[2].record()
            img_ = _ops.weight_image(w.view(K, 1, CIN, COUT), K, 1, CIN, COUT, False)
            ev[3].record()
            _ops.gather_gemm(x, img_, plan_, 1, CIN, COUT)
            ev[4].record()
            sparse_conv_dgrad(gy, w, km_, n)
            ev[5].record()
            sparse_conv_wgrad(x, gy, (K, CIN, COUT), km_)
            ev[6].record()
            torch.cuda.synchronize()
            acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))])
        return {k: round(float(v / reps), 4) for k, v in zip(names, acc)}

    phases = breakdown()
    phases["note"] = ("eager launches with events between phases: each entry is max(host enqueue "
                      "time, GPU time) of the phase")

    # ---- end to end through the public API from pinned host buffers ---------------------------
    conv = SparseConv3d(CIN, COUT, KS, bias=False).to(dev)
    with torch.no_grad():
        conv.weight.copy_(w_h)
    coords_pin = torch.from_numpy(coords).pin_memory()
    feats_pin = x_h.bfloat16().pin_memory()
    dw_pin = torch.empty((K, CIN, COUT), dtype=torch.float32).pin_memory()
    offsets = torch.tensor([0, n], dtype=torch.int64)
    h2d = coords_pin.numel() * 4 + feats_pin.numel() * 2
    d2h = dw_pin.numel() * 4

    # Double-buffered pipeline: the H2D copy of step i+1 runs on a copy stream while step i
    # computes; every step's inputs are still copied from pinned host memory inside the timed
    # region (the first copy is ordered after the start event), and every step ends with the D2H
    # read of its weight gradient.
    copy_stream = torch.cuda.Stream(device=dev)
    cbuf = [torch.empty((n, 3), dtype=torch.int32, device=dev) for _ in range(2)]
    fbuf = [torch.empty((n, CIN), dtype=torch.bfloat16, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i % 2])
            cbuf[i % 2].copy_(coords_pin, non_blocking=True)
            fbuf[i % 2].copy_(feats_pin, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_run(steps):
        cur = torch.cuda.current_stream()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for ev in freed:
            ev.record(cur)
        s.record(cur)
        copy_stream.wait_event(s)
        prefetch(0)
        for i in range(steps):
            if i + 1 < steps:
                prefetch(i + 1)
            cur.wait_event(ready[i % 2])
            f = fbuf[i % 2].detach().requires_grad_(True)
            vox = Voxels(cbuf[i % 2], f, offsets=offsets)
            conv.weight.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = conv(vox)
            out.feature_tensor.backward(gy)
            freed[i % 2].record(cur)
            g = conv.weight.grad
            if par is not None:
                par.buffer.copy_(g.reshape(-1))
                g = par.all_reduce_().view(K, CIN, COUT)
            elif world > 1:
                dist.all_reduce(g)
            dw_pin.copy_(g, non_blocking=True)
        e.record(cur)
        return s, e

    # untimed warm-up of the e2e loop: on a freshly booted box the first process sees ~1.5x slower
    # host->device copies for its first few dozen steps (link / host clocks ramping up; a second
    # process on the same box does not), so the loop is run for 50 steps before the K timed ones
    if sampler is not None:
        # the e2e loop is paced by the host thread: sample clocks every 20 ms there so the sampler
        # thread does not compete with it for the interpreter lock
        sampler.period = 0.02
    e2e_run(max(args.warmup, 50))
    barrier()
    # The loop is paced by the host thread and the PCIe copies, so single K-step measurements
    # scatter (1.0 - 1.4 ms on the same box): K steps are timed five times, each max-reduced over
    # ranks, and the MEDIAN repetition is reported (all three are listed in e2e.runs_ms).
    e2e_runs = []
    for _ in range(5):
        s_ev, e_ev = e2e_run(args.steps)
        barrier()
        t = torch.tensor([s_ev.elapsed_time(e_ev) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_runs.append(float(t.item()))
    e2e_ms = float(np.median(e2e_runs))
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = ("all timed regions of this run (graph steps, eager steps, kernel-only "
                            "launches, e2e loop)")

    # ---- side blocks (never allowed to take the headline line down) ----------------------------
    side = {}
    if world == 1 and not args.no_side:
        try:
            side = run_side_blocks(dev, flush)
        except Exception as exc:  # pragma: no cover
            side = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
    ref_gpu = None
    if world == 1 and rank == 0 and args.ref_gpu != "none":
        torch.cuda.empty_cache()
        ref_gpu = run_ref_gpu(args.ref_gpu.split(","), args.ref_gpu_timeout)
    c4 = None
    if not args.no_c4:
        # free the C3 working set first
        try:
            c4 = run_c4(dev, rank, world, dist, steps=max(5, min(args.steps, 10)), warmup=3)
        except Exception as exc:  # pragma: no cover
            import traceback
            c4 = {"error": f"{type(exc).__name__}: {str(exc)[:300]}",
                  "trace": traceback.format_exc()[-800:]}
            torch.cuda.synchronize()

    c5 = None
    if not args.no_c4:
        try:
            c5 = run_c5(dev, rank, world, dist, flush)
        except Exception as exc:  # pragma: no cover
            import traceback
            c5 = {"error": f"{type(exc).__name__}: {str(exc)[:300]}",
                  "trace": traceback.format_exc()[-800:]}
            torch.cuda.synchronize()

    peaks = load_peaks()
    steps_total = int(plan.tile_nk.sum().item())
    # Secondary (informational) bound, DESIGN.md 4.3: every gathered row crosses the L2->SM path
    # through the LSU (cp.async), which tools/l2sm_bench.cu measures at 26.9 B/cycle/SM = 6.1 TB/s
    # chip-wide for 256-byte row gathers on this pool's B200 (sequential or random rows alike,
    # TMA gather4 slower). The per-step weight slices (0.88 MB image, hot in L2, cp.async.bulk)
    # ride on the TMA path and do not compete (same benchmark, "GB" rows).
    gather_bytes = L * CIN * 2
    l2sm = {"gathered_bytes_per_launch": gather_bytes,
            "weight_slice_bytes_per_launch": steps_total * CIN * COUT * 2,
            "achieved_TBps": gather_bytes / (gemm_ms * 1e-3) / 1e12,
            "measured_gather_cap_TBps": 6.1,
            "cap_source": "tools/l2sm_bench.cu, mode G at 148 CTAs (gpurun_out -> "
                          "profiles/r1c_gather_path.md)",
            "frac_of_cap": gather_bytes / (gemm_ms * 1e-3) / 1e12 / 6.1}
    flops = 2.0 * L * CIN * COUT
    achieved = flops / (gemm_ms * 1e-3) / 1e12
    out = {
        "metric": "SparseConv3d fwd+bwd voxels/sec", "value": total_vox / (ms * 1e-3),
        "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        # `config` is identical in both arms (--impl ours / reference); what differs is in `setup`
        "config": {"workload": workload_name(args.dist, n, world), "voxels_per_gpu": n},
        "setup": {
            "pairs_L": L, "parallelism": f"scene-sharded dp{world}",
            "collective": {"none": "none (one rank)",
                           "peer": "own kernel over NVLink peer memory (wcn_peer_allreduce_f32), "
                                   "side stream, under dgrad"}.get(collective, collective),
            "l2": "flushed with a 256 MiB write before every timed step (outside the events)",
            "timing": "CUDA events per step on the launching stream, mean over steps, max over ranks",
            "launch": ("one CUDA-graph replay per step (the path has no host sync)" if graph is not None
                       else (graph_note or "eager launches")),
            "eager_ms_per_step": eager_ms,
            "peer_allreduce_barrier_timeouts": None if par is None else par.timeouts(),
        },
        "e2e": {"value": total_vox / (e2e_ms * 1e-3), "unit": "voxels/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "h2d_GBps_per_rank": h2d / (e2e_ms * 1e-3) / 1e9, "numa_binding": numa,
                "runs_ms": e2e_runs, "runs": "5 repetitions of K steps, median reported",
                "warmup_steps": max(args.warmup, 50),
                "api": "Voxels(pinned host coords+feats) -> SparseConv3d.forward (autocast bf16) -> "
                       "backward -> weight.grad to pinned host",
                "pipeline": "H2D of step i+1 overlaps compute of step i (copy stream, 2 buffers); "
                            "K steps bracketed by one event pair; working set 250 MB > L2, inputs "
                            "re-copied from host every step, no explicit flush in this loop"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        # the forward kernel is timed ALONE (one launch between L2 flushes, SM clock at its
        # maximum), so the denominator is the BURST bf16 peak; the fraction of the sustained peak
        # (what a long dense-matmul step reaches on this part) is given next to it
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["burst"],
                     "unit": "TFLOP/s", "frac": achieved / peaks["burst"],
                     "peak_kind": "burst bf16 dense matmul (kernel timed alone)",
                     "frac_of_sustained_peak": achieved / peaks["tflops"],
                     "sustained_peak": peaks["tflops"],
                     # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload
                     # from this round's `ncu --set full` capture of the shipped kernel (71.1 +
                     # 24.0 MB per launch, profiles/r2p_ncu_full_fwd_wgrad.md); algorithmic HBM
                     # bytes: 111 MB
                     "traffic": 95.1e6 if args.dist == "S" else None,
                     "traffic_unit": "bytes per launch (ncu --set full capture r2p, C3-S)",
                     "kernel": "gather_gemm_kernel<bf16> (forward AB_gather_scatter)",
                     "kernel_ms": gemm_ms, "flops_per_launch": flops, "peak_source": peaks["which"],
                     # what actually bounds the kernel (profiles/): bytes that must
                     # cross the L2->SM crossbar = gathered rows (re-fetched once per offset that
                     # uses them) + weight slices (once per step of a 256-row tile) + step indices
                     "l2_to_sm": l2sm},
        "phases_ms": phases,
        "wall_s_timed_region": wall,
        "c4": c4,
        "c5": c5,
        "ref_gpu": ref_gpu,
    }
    out.update(side)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"], _ = run_cpu_baseline(args.dist, reps=2, warmup=1)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        # Tear-down: release the captured graph (it holds NCCL work) before the communicator.
        # ncclCommDestroy has been seen to block forever after a graph-captured all-reduce on this
        # stack, which would stall the launcher; every rank therefore synchronises, meets at a
        # barrier and leaves through os._exit(0) once its output is flushed.
        graph = None
        g_out = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    base, n = run_cpu_baseline(args.dist, reps=max(1, args.steps), warmup=min(args.warmup, 1),
                               budget_s=150.0)
    ms = base["seconds_per_pass"] * 1e3
    out = {
        "impl": "reference", "metric": "SparseConv3d fwd+bwd voxels/sec", "value": base["value"],
        "unit": "voxels/s", "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same workload name as our arm's line; what differs is said in `arm`
        "config": {"workload": workload_name(args.dist, n, int(os.environ.get("WORLD_SIZE", 1))),
                   "voxels_per_gpu": n},
        "setup": {"arm": "host CPU, fp32: oracle kernel map + port of the reference's explicit "
                         "gather-matmul-scatter (detail/explicit.py:22-101), rank 0 only; the "
                         "reference's own GPU build is timed in the `ref_gpu` block of the other arm"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--collective", choices=["peer", "nccl"], default="peer",
                    help="all-reduce of dW at N > 1: own peer-memory kernel (default) or NCCL")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dist", choices=["S", "R"], default="S")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph replay")
    ap.add_argument("--no-c4", action="store_true", help="skip the MinkUNet-14 (config C4) and C5 blocks")
    ap.add_argument("--no-side", action="store_true", help="skip the C3-R / C2 / C5 side blocks")
    ap.add_argument("--ref-gpu", default="kmap,c3s", help="sections of tools/ref_gpu_bench.py to run "
                    "against the built reference (N = 1 only): kmap,c3s,c3r,c4 or 'none'; c4 needs "
                    "~5 min for the reference's auto-tuner")
    ap.add_argument("--ref-gpu-timeout", type=int, default=240)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
