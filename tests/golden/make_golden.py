# SPDX-License-Identifier: Apache-2.0
"""Generates the committed golden fixtures of tests/golden/*.npz.  TEST INFRASTRUCTURE ONLY.

Runs only where the reference tree is mounted (/root/reference, i.e. the build container — never
on the GPU box). For every case it stores

* the inputs (batch-indexed coordinates, features, weights, output gradients),
* the kernel map in three forms: the brute-force Python-dict enumeration the reference's own
  tests use as their pin (tests/coords/test_kernel_map_invariants.py:181-277) as sorted
  (k, in, out) triples, plus ``offsets`` / ``pair_table`` / CSR from ``oracle.kernel_map``
  (asserted equal to the dict enumeration before anything is written),
* the outputs of the REFERENCE's own ``_explicit_gemm_forward_logic`` /
  ``_explicit_gemm_backward_logic`` (warpconvnet/nn/functional/sparse_conv/detail/explicit.py:
  22-101) run in fp32 and fp64 on CPU on that map, wrapped in the reference's IntSearchResult.

Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import kernel_map as okm  # noqa: E402
from oracle import ref_adapter  # noqa: E402


def c1_coords():
    """1000 unique voxels of a 16^3 grid, randperm(4096, seed 0)[:1000]."""
    g = torch.Generator().manual_seed(0)
    idx = torch.randperm(4096, generator=g)[:1000].numpy()
    c = np.stack([idx // 256, (idx // 16) % 16, idx % 16], axis=1).astype(np.int32)
    return okm.batch_indexed([c])


def toy_coords():
    """The reference's `toy_voxels` fixture (tests/conftest.py:148-188): 2 scenes x 7 voxels."""
    b0 = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [2, 2, 2], [3, 1, 3], [4, 3, 0]]
    b1 = [[4, 4, 4], [5, 4, 4], [4, 5, 4], [4, 4, 5], [6, 2, 0], [3, 6, 1], [7, 1, 3]]
    return okm.batch_indexed([np.array(b0, np.int32), np.array(b1, np.int32)])


def run_reference(x, w, gy, km, n_out, dtype):
    fwd, bwd, ISR = ref_adapter.load()
    ref_km = ISR(torch.from_numpy(km["in_maps"]).long(), torch.from_numpy(km["out_maps"]).long(),
                 torch.from_numpy(km["offsets"]).long(),
                 identity_map_index=km["identity_map_index"])
    xt, wt, gt = (torch.from_numpy(a).to(dtype) for a in (x, w, gy))
    y = fwd(xt, wt, ref_km, n_out)
    dx, dw = bwd(gt, xt, wt, ref_km)
    return y.numpy(), dx.numpy(), dw.numpy()


def make_case(name, in_bc, stride, ksize, cin, cout, seed):
    if all(s == 1 for s in stride):
        out_bc = in_bc
    else:
        out_bc, _ = okm.stride_coords(in_bc, stride)
    km = okm.generate_kernel_map(in_bc, out_bc, stride, ksize)
    brute = okm.brute_force_pairs(in_bc, out_bc, stride, ksize)
    got = set()
    for k in range(len(km["offsets"]) - 1):
        s, e = km["offsets"][k], km["offsets"][k + 1]
        got |= {(k, int(i), int(o)) for i, o in zip(km["in_maps"][s:e], km["out_maps"][s:e])}
    assert got == brute, f"{name}: oracle kernel map != brute-force dict enumeration"
    triples = np.array(sorted(brute), dtype=np.int32).reshape(-1, 3)
    K = int(np.prod(ksize))
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(len(in_bc), cin, generator=g).numpy()
    w = (torch.randn(K, cin, cout, generator=g) * (K * cin) ** -0.5).numpy()
    gy = torch.randn(len(out_bc), cout, generator=g).numpy()
    y32, dx32, dw32 = run_reference(x, w, gy, km, len(out_bc), torch.float32)
    y64, dx64, dw64 = run_reference(x, w, gy, km, len(out_bc), torch.float64)
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(
        path, in_bcoords=in_bc, out_bcoords=out_bc, stride=np.array(stride, np.int32),
        kernel_size=np.array(ksize, np.int32), triples=triples, offsets=km["offsets"],
        pair_table=km["pair_table"], in_maps=km["in_maps"], out_maps=km["out_maps"],
        identity_map_index=np.int32(-1 if km["identity_map_index"] is None
                                    else km["identity_map_index"]),
        x=x, w=w, gy=gy, y_ref_f32=y32, dx_ref_f32=dx32, dw_ref_f32=dw32, y_ref_f64=y64,
        dx_ref_f64=dx64, dw_ref_f64=dw64)
    print(f"wrote {path}: N={len(in_bc)} M={len(out_bc)} L={int(km['offsets'][-1])} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    if not ref_adapter.available():
        raise SystemExit("reference tree not mounted; fixtures can only be generated in the "
                         "build container")
    make_case("c1_s1_k3", c1_coords(), (1, 1, 1), (3, 3, 3), 4, 8, seed=1)     # BASELINE config 1
    make_case("c1_s2_k2", c1_coords(), (2, 2, 2), (2, 2, 2), 4, 8, seed=2)     # strided down conv
    make_case("c1_s2_k3", c1_coords(), (2, 2, 2), (3, 3, 3), 8, 4, seed=3)     # strided 3^3
    make_case("toy_s1_k3", toy_coords(), (1, 1, 1), (3, 3, 3), 3, 5, seed=4)   # 2 scenes
    make_case("toy_s2_k2", toy_coords(), (2, 2, 2), (2, 2, 2), 3, 5, seed=5)
    make_case("c1_s1_k5", c1_coords(), (1, 1, 1), (5, 5, 5), 4, 4, seed=6)     # K=125


if __name__ == "__main__":
    main()
