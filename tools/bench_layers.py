# SPDX-License-Identifier: Apache-2.0
"""Per-layer table of the three sparse-conv GEMMs inside one MinkUNet-14 (config C4) step:
every call of wcn_gather_gemm / wcn_wgrad is bracketed by CUDA events (eager launches), and listed
with its shape, pair count, algorithmic FLOP/s and the bytes it has to gather through the LSU path
(B/cycle/SM against the measured 32 B/cycle/SM limit of that path).
python tools/bench_layers.py [scenes] > gpurun_out/layers.md"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from minkunet14 import MinkUNet14, surface_scene  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
from warpconvnet_b200.geometry.types.voxels import Voxels  # noqa: E402

scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SM, GHZ = 148, 1.92
coords = [surface_scene(548, s).cuda() for s in range(scenes)]
feats = [torch.randn(len(c), 3, device="cuda") for c in coords]
net = MinkUNet14(3, 20).cuda()
log = []
real_gg, real_wg = _ops.gather_gemm, _ops.wgrad


def timed(kind, fn, meta):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(400000)  # ~200 us of GPU spin: the host enqueues event, kernel(s), event behind it
    a.record()
    out = fn()
    b.record()
    log.append((kind, meta, a, b))
    return out


def gg(feats_, wimg, plan, groups, cin_g, cout_g, *args, **kw):
    meta = dict(rows_out=plan.n_rows, rows_in=feats_.shape[0], cin=groups * cin_g, cout=groups * cout_g,
                K=plan.K, plan=plan, kflip=bool(kw.get("kflip", False)), stats=kw.get("stats") is not None)
    return timed("gemm", lambda: real_gg(feats_, wimg, plan, groups, cin_g, cout_g, *args, **kw), meta)


def wg(feats_, gout, in_maps, out_maps, offsets_dev, K, groups, cin_g, cout_g, *args, **kw):
    meta = dict(rows_out=gout.shape[0], rows_in=feats_.shape[0], cin=groups * cin_g, cout=groups * cout_g,
                K=K, offsets=offsets_dev)
    return timed("wgrad", lambda: real_wg(feats_, gout, in_maps, out_maps, offsets_dev, K, groups,
                                           cin_g, cout_g, *args, **kw), meta)


real_bnf, real_bnb, real_ssa = _ops.bn_forward, _ops.bn_backward, _ops.scale_shift_act
in_bnf = [False]


def bnf(x, gamma, beta, eps, momentum, rm, rv, residual, relu, sums=None):
    n, c = x.shape
    passes = (1 if sums is not None else 2) + (1 if residual is not None else 0) + 1
    meta = dict(n=n, c=c, what="bn fwd" + ("" if sums is not None else " +stats pass") +
                (" +res" if residual is not None else "") + (" +relu" if relu else ""),
                bytes=passes * n * c * x.element_size())
    in_bnf[0] = True
    try:
        return timed("norm", lambda: real_bnf(x, gamma, beta, eps, momentum, rm, rv, residual, relu, sums=sums), meta)
    finally:
        in_bnf[0] = False


def ssa(x, scale, shift, residual=None, relu=False):
    if in_bnf[0]:
        return real_ssa(x, scale, shift, residual, relu)
    n, c = x.shape
    meta = dict(n=n, c=c, what="scale_shift_act", bytes=(2 + (residual is not None)) * n * c * x.element_size())
    return timed("norm", lambda: real_ssa(x, scale, shift, residual, relu), meta)


def bnb(dy, x, y, gamma, mean_rstd, msc, msh, want_dres):
    n, c = x.shape
    per = 2 + (y is not None)                       # reduce: dy, x (, y)
    passes = per + per + 1 + (1 if want_dres else 0)  # apply: same reads + dx (+ dres)
    meta = dict(n=n, c=c, what="bn bwd (reduce + apply)" + (" mask from y" if y is not None else
                (" mask from x" if msc is not None else "")) + (" +dres" if want_dres else ""),
                bytes=passes * n * c * x.element_size())
    return timed("norm", lambda: real_bnb(dy, x, y, gamma, mean_rstd, msc, msh, want_dres), meta)


def step():
    x = Voxels(coords, feats)
    net.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(x)
    out.feature_tensor.float().square().mean().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
_ops.gather_gemm, _ops.wgrad = gg, wg
_ops.bn_forward, _ops.bn_backward, _ops.scale_shift_act = bnf, bnb, ssa
import warpconvnet_b200.nn.functional.sparse_conv.detail.unified as uni  # noqa: E402
assert uni._ops is _ops
REPS = 3
runs = []
for _ in range(REPS):
    log.clear()
    step()
    torch.cuda.synchronize()
    runs.append([(k, m, a.elapsed_time(b) * 1e3) for k, m, a, b in log])
_ops.gather_gemm, _ops.wgrad = real_gg, real_wg
_ops.bn_forward, _ops.bn_backward, _ops.scale_shift_act = real_bnf, real_bnb, real_ssa

print(f"# MinkUNet-14 (C4), {scenes} scenes = {sum(len(c) for c in coords)} voxels: per-call table of the sparse-conv GEMMs")
print()
print("eager launches, CUDA events around each call, enqueued behind a 200 us GPU spin so that the events "
      "bracket the kernel and not the host's launch latency; best of 3 steps; `pairs` = (in, out) pairs of the kernel map, `slots` "
      "= rows x steps of the tile plan (forward / dgrad gather one row per slot, zero rows included); "
      "gather B/clk/SM = gathered bytes / time / 148 SMs / 1.92 GHz against the 32 B/clk/SM limit of the "
      "LSU gather path; TFLOP/s algorithmic (2 x pairs x cin x cout).")
print()
print("| # | call | rows in -> out | cin -> cout | K | pairs | slots/pairs | us | TFLOP/s | gather B/clk/SM |")
print("|---:|---|---|---|---:|---:|---:|---:|---:|---:|")
tot = {"gemm": 0.0, "wgrad": 0.0, "norm": 0.0}
norm_rows = []
for i in range(len(runs[0])):
    kind, m, _ = runs[0][i]
    us = min(r[i][2] for r in runs)
    tot[kind] += us
    if kind == "norm":
        norm_rows.append((i, m, us))
        continue
    if kind == "gemm":
        plan = m["plan"]
        nk = plan.tile_nk[:plan.num_tiles].sum().item()
        slots = nk * plan.tile_rows
        # real pairs: neighbours >= 0 in the step lists
        pairs = None
        waste = ""
        gathered = slots * m["cin"] * 2
        flops_pairs = slots  # upper bound; replaced below when cheap to count
        if plan.step_nbr.numel() <= (1 << 28):
            valid = 0
            sn = plan.step_nbr.view(-1, plan.K, plan.tile_rows)
            kk = torch.arange(plan.K, device=sn.device)[None, :, None]
            mask = kk < plan.tile_nk[:sn.shape[0], None, None]
            valid = int(((sn >= 0) & mask).sum().item())
            pairs = valid
            waste = f"{slots / max(valid, 1):.2f}"
            gathered = valid * m["cin"] * 2
        name = ("dgrad" if m["kflip"] or m["rows_out"] != plan.n_rows else "fwd/dgrad") + (" +stats" if m["stats"] else "")
        fl = 2.0 * (pairs or slots) * m["cin"] * m["cout"]
    else:
        pairs = int(m["offsets"][-1].item())
        waste = "1.00"
        gathered = pairs * (m["cin"] + m["cout"]) * 2
        name = "wgrad"
        fl = 2.0 * pairs * m["cin"] * m["cout"]
    bpc = gathered / (us * 1e-6) / SM / (GHZ * 1e9)
    print(f"| {i} | {name} | {m['rows_in']} -> {m['rows_out']} | {m['cin']} -> {m['cout']} | {m['K']} | {pairs} | "
          f"{waste} | {us:.1f} | {fl / (us * 1e-6) / 1e12:.1f} | {bpc:.1f} |")
print()
print(f"sum: gather-GEMM calls {tot['gemm'] / 1e3:.2f} ms, wgrad calls {tot['wgrad'] / 1e3:.2f} ms per step")
print()
print("## BatchNorm / ReLU / residual tail (HBM-bound: bytes = full passes over the [n, c] tensor)")
print()
print("| # | call | n x c | MB moved | us | TB/s |")
print("|---:|---|---|---:|---:|---:|")
tb = 0
for i, m, us in norm_rows:
    tb += m["bytes"]
    print(f"| {i} | {m['what']} | {m['n']} x {m['c']} | {m['bytes'] / 1e6:.1f} | {us:.1f} | {m['bytes'] / (us * 1e-6) / 1e12:.2f} |")
print()
print(f"sum: norm calls {tot['norm'] / 1e3:.2f} ms per step, {tb / 1e9:.2f} GB moved = {tb / (tot['norm'] * 1e-6) / 1e12:.2f} TB/s average")
