# SPDX-License-Identifier: Apache-2.0
"""Per-role wait-cycle counters of the gather-GEMM kernel (bring-up only; needs a library built with
WCN_KERNEL_COUNTERS=1 warpconvnet_b200/csrc/build.sh)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import surface_coords  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map  # noqa: E402

cin = cout = 128
c = surface_coords(448, 0)
n = len(c)
bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), c], 1)).cuda()
km = generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), same_coords=True)
x = torch.randn(n, cin, device="cuda").bfloat16()
w = (torch.randn(27, 1, cin, cout, device="cuda") * 0.02).bfloat16()
img = _ops.weight_image(w, 27, 1, cin, cout, False)
table = km.pair_table(n)
dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
os.environ["WCN_DEBUG_PTR"] = str(dbg.data_ptr())
names = ["prod_total", "prod_wait_empty", "mma_total", "mma_wait_full", "mma_wait_accempty",
         "epi_total", "epi_wait_accfull", "prod_ns"]
for tr in (256,):
    plan = _ops.build_tile_plan(table, tile_rows=tr)
    for flags in (0,):
        os.environ["WCN_DEBUG"] = str(flags)
        for _ in range(3):
            dbg.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _ops.gather_gemm(x, img, plan, 1, cin, cout)
            b.record()
            torch.cuda.synchronize()
        d = dbg.view(148, 16).cpu().numpy()
        print(f"tile_rows={tr} dbg={flags} time={a.elapsed_time(b) * 1e3:.1f} us  steps={int(plan.tile_nk.sum())}")
        print(f"    SM clock (prod cycles / ns): {(d[:, 0] / d[:, 7]).mean():.3f} GHz")
        for i, nm in enumerate(names):
            print(f"    {nm:18s} mean={d[:, i].mean():10.0f} min={d[:, i].min():10d} max={d[:, i].max():10d}")
        t0 = d[:, 8].min()
        print(f"    ns since first CTA entry: entry mean={np.mean(d[:, 8] - t0):.0f} max={np.max(d[:, 8] - t0):.0f} | "
              f"prod start mean={np.mean(d[:, 9] - t0):.0f} max={np.max(d[:, 9] - t0):.0f} | "
              f"prod end mean={np.mean(d[:, 10] - t0):.0f} max={np.max(d[:, 10] - t0):.0f} | "
              f"all roles done mean={np.mean(d[:, 11] - t0):.0f} max={np.max(d[:, 11] - t0):.0f}")
        if flags == 0:
            np.set_printoptions(linewidth=200)
            print("    prod_total per CTA (kcycles):")
            print((d[:, 0] // 1000).reshape(4, 37))
            print("    prod_wait_empty per CTA (kcycles):")
            print((d[:, 1] // 1000).reshape(4, 37))
            print("    mma_wait_full per CTA (kcycles):")
            print((d[:, 3] // 1000).reshape(4, 37))

# ---- per-CTA work composition vs measured producer time (static ranges, unit = 128 rows) ---------
plan = _ops.build_tile_plan(table, tile_rows=256)
os.environ["WCN_DEBUG"] = "0"
dbg.zero_()
_ops.gather_gemm(x, img, plan, 1, cin, cout)
torch.cuda.synchronize()
d = dbg.view(148, 16).cpu().numpy()
nk = plan.tile_nk.cpu().numpy().astype(np.int64)
cum = plan.tile_cum.cpu().numpy().astype(np.int64)
nt = plan.num_tiles
U = 2
S = U * cum[nt]
unit_cost = np.zeros(U * nt + 1, np.int64)
for u in range(U * nt + 1):
    t = min(u // U, nt - 1) if u < U * nt else nt
    unit_cost[u] = U * cum[nt] if u == U * nt else U * cum[t] + (u - t * U) * nk[t]
G = 148
ub = [int(np.searchsorted(unit_cost, S * b // G, side="left")) for b in range(G + 1)]
if plan.cta_units is not None and plan.n_range_ctas == G:
    ub = plan.cta_units.cpu().numpy().astype(np.int64).tolist()   # the ranges the kernel really used
nbr = plan.step_nbr.cpu().numpy().reshape(nt, plan.K, 256)
valid_per_unit = np.zeros(U * nt, np.int64)
for t in range(nt):
    v = (nbr[t, :nk[t]] >= 0)
    valid_per_unit[2 * t] = v[:, :128].sum()
    valid_per_unit[2 * t + 1] = v[:, 128:].sum()
rows = []
for b in range(G):
    u0, u1 = ub[b], ub[b + 1]
    steps = sum(nk[u // U] for u in range(u0, u1))
    valid = valid_per_unit[u0:u1].sum()
    partial = (u0 % 2) + (u1 % 2)
    rows.append((b, u1 - u0, steps, valid, partial, d[b, 0]))
np.save(os.path.join(ROOT, "gpurun_out", "dbg_rows.npy"), np.array(rows))
np.save(os.path.join(ROOT, "gpurun_out", "dbg_raw.npy"), d)
rows = np.array(rows)
print("corr(prod_total, unit-steps) =", np.corrcoef(rows[:, 5], rows[:, 2])[0, 1])
print("corr(prod_total, valid rows) =", np.corrcoef(rows[:, 5], rows[:, 3])[0, 1])
print("corr(prod_total, partial tiles) =", np.corrcoef(rows[:, 5], rows[:, 4])[0, 1])
order = np.argsort(-rows[:, 5])
print("slowest CTAs: (cta, units, unit-steps, valid rows, partial tiles, prod kcycles)")
for i in order[:8]:
    print("   ", rows[i, :5].tolist(), rows[i, 5] // 1000)
print("fastest CTAs:")
for i in order[-8:]:
    print("   ", rows[i, :5].tolist(), rows[i, 5] // 1000)
# least-squares model of a CTA's time: a * unit-steps + b * valid rows + c * partial tiles + d
A = np.stack([rows[:, 2], rows[:, 3], rows[:, 4], np.ones(len(rows))], 1).astype(np.float64)
for tgt, nm in ((rows[:, 5].astype(np.float64), "prod_total"), ((d[:, 11] - d[:, 8]).astype(np.float64), "cta_ns")):
    coef, res, _, _ = np.linalg.lstsq(A, tgt, rcond=None)
    pred = A @ coef
    print(f"fit {nm}: per unit-step {coef[0]:.2f}, per valid row {coef[1]:.4f}, per partial tile {coef[2]:.1f}, "
          f"const {coef[3]:.0f}; rms residual {np.sqrt(np.mean((tgt - pred) ** 2)):.0f} of mean {tgt.mean():.0f} "
          f"(spread: min {tgt.min():.0f} max {tgt.max():.0f})")
print("unit-steps per CTA: min", rows[:, 2].min(), "max", rows[:, 2].max(), "| valid rows per CTA: min",
      rows[:, 3].min(), "max", rows[:, 3].max())
