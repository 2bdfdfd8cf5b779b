# SPDX-License-Identifier: Apache-2.0
"""Host-side timeline of bench.py's e2e loop: where does the host thread block when a repetition
takes 3-6 ms per step instead of 1.0? Records perf_counter around every call of every step."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from warpconvnet_b200.geometry.types.voxels import Voxels  # noqa: E402
from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d  # noqa: E402

K, CIN, COUT, KS = bench.K, bench.CIN, bench.COUT, bench.KS
dev = torch.device("cuda", 0)
coords = bench.make_coords("S", seed=0)
n = len(coords)
x_h, w_h, gy_h = bench.make_tensors(n, seed=0)
gy = gy_h.to(dev).bfloat16()
conv = SparseConv3d(CIN, COUT, KS, bias=False).to(dev)
coords_pin = torch.from_numpy(coords).pin_memory()
feats_pin = x_h.bfloat16().pin_memory()
dw_pin = torch.empty((K, CIN, COUT), dtype=torch.float32).pin_memory()
offsets = torch.tensor([0, n], dtype=torch.int64)
copy_stream = torch.cuda.Stream(device=dev)
cbuf = [torch.empty((n, 3), dtype=torch.int32, device=dev) for _ in range(2)]
fbuf = [torch.empty((n, CIN), dtype=torch.bfloat16, device=dev) for _ in range(2)]
ready = [torch.cuda.Event() for _ in range(2)]
freed = [torch.cuda.Event() for _ in range(2)]
NAMES = ["prefetch", "wait+voxels", "forward", "backward", "d2h"]


# ---- fine-grained host timers inside forward: every _ops entry point + the map builder ----------
import warpconvnet_b200._ops as _ops  # noqa: E402
import warpconvnet_b200.geometry.coords.search.torch_discrete as td  # noqa: E402
import warpconvnet_b200.geometry.coords.search.packed_hashmap as ph  # noqa: E402
FINE = {}


def _wrap(mod, name):
    fn = getattr(mod, name)

    def timed_fn(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            FINE[name] = FINE.get(name, 0.0) + (time.perf_counter() - t) * 1e3
    setattr(mod, name, timed_fn)


for nm in ("kernel_map_search_symmetric", "kernel_map_count", "kernel_map_scatter", "build_tile_plan",
           "weight_image_pair", "gather_gemm", "wgrad", "hash_insert", "hash_prepare"):
    if hasattr(_ops, nm):
        _wrap(_ops, nm)
_wrap(td, "check_pending_kernel_maps")
_wrap(torch, "cat")
_wrap(torch, "empty")
_wrap(torch, "zeros")


def prefetch(i):
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(freed[i % 2])
        cbuf[i % 2].copy_(coords_pin, non_blocking=True)
        fbuf[i % 2].copy_(feats_pin, non_blocking=True)
        ready[i % 2].record(copy_stream)


def run(steps):
    cur = torch.cuda.current_stream()
    for ev in freed:
        ev.record(cur)
    prefetch(0)
    log = np.zeros((steps, len(NAMES)))
    fine_log = []
    allocs = []
    for i in range(steps):
        FINE.clear()
        a0 = torch.cuda.memory_stats()["num_device_alloc"]
        t0 = time.perf_counter()
        if i + 1 < steps:
            prefetch(i + 1)
        t1 = time.perf_counter()
        cur.wait_event(ready[i % 2])
        f = fbuf[i % 2].detach().requires_grad_(True)
        vox = Voxels(cbuf[i % 2], f, offsets=offsets)
        conv.weight.grad = None
        t2 = time.perf_counter()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = conv(vox)
        t3 = time.perf_counter()
        out.feature_tensor.backward(gy)
        t4 = time.perf_counter()
        freed[i % 2].record(cur)
        dw_pin.copy_(conv.weight.grad, non_blocking=True)
        t5 = time.perf_counter()
        log[i] = [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4]
        fine_log.append(dict(FINE))
        allocs.append(torch.cuda.memory_stats()["num_device_alloc"] - a0)
    torch.cuda.synchronize()
    run.fine, run.allocs = fine_log, allocs
    return log * 1e3


import gc  # noqa: E402
run(50)
for rep in range(16):
    if rep == 8:
        gc.collect()
        gc.disable()
        print("---- cyclic GC disabled from here")
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    t0 = time.perf_counter()
    log = run(40)
    wall = (time.perf_counter() - t0) * 1e3
    e.record()
    torch.cuda.synchronize()
    tot = log.sum(1)
    print(f"rep {rep}: gc counts {gc.get_count()} {s.elapsed_time(e) / 40:.3f} ms/step (events), host wall {wall / 40:.3f} ms/step, "
          f"host per-step median {np.median(tot):.3f} max {tot.max():.3f}")
    for i in np.argsort(-tot)[:3]:
        if tot[i] > 1.5:
            print("    step", i, {k: round(float(v), 3) for k, v in zip(NAMES, log[i])},
                  "cudaMallocs", run.allocs[i],
                  {k: round(v, 3) for k, v in sorted(run.fine[i].items(), key=lambda kv: -kv[1])[:4]})
print("pinned host allocator:", {k: v for k, v in torch.cuda.host_memory_stats().items()
                                   if "num_host_alloc" in k or "allocated_bytes.current" in k or k.startswith("host_alloc_time")}
      if hasattr(torch.cuda, "host_memory_stats") else "n/a")
