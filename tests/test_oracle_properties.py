# SPDX-License-Identifier: Apache-2.0
"""CPU property tests of the oracle itself (no GPU): the vectorised kernel-map restatement against
the Python-dictionary enumeration the reference's own tests use as their pin
(tests/coords/test_kernel_map_invariants.py:181-277) over a seeded sweep of shapes, the reverse
table, strided / expanded coordinate sets, and the design claim behind the 24-bit mask sort keys
(csrc/cuhash.cu) evaluated on the oracle's masks."""
import numpy as np
import pytest

from conftest import random_coords, surface_coords
from oracle import kernel_map as okm


def _draws(n_draws, seed):
    rng = np.random.RandomState(seed)
    for it in range(n_draws):
        ks = tuple(int(k) for k in rng.choice([1, 2, 3, 4], size=3))
        stride = tuple(int(s) for s in rng.choice([1, 1, 2, 3], size=3))
        dil = tuple(int(d) for d in rng.choice([1, 1, 2], size=3))
        sizes = [int(s) for s in rng.randint(1, 160, size=rng.randint(1, 4))]
        shift = rng.randint(-30, 5, size=3).astype(np.int32)
        bc = okm.batch_indexed([random_coords(n, float(rng.uniform(0.1, 0.6)), it * 5 + j) + shift
                                for j, n in enumerate(sizes)])
        yield it, ks, stride, dil, bc


def test_pair_table_equals_dictionary_enumeration_over_a_shape_sweep():
    for it, ks, stride, dil, bc in _draws(40, 7):
        out_bc = bc if stride == (1, 1, 1) else okm.stride_coords(bc, stride)[0]
        km = okm.generate_kernel_map(bc, out_bc, stride, ks, dilation=dil)
        pairs = okm.brute_force_pairs(bc, out_bc, stride, ks, dil)
        got = set()
        offs = km["offsets"]
        for k in range(len(offs) - 1):
            for i, o in zip(km["in_maps"][offs[k]:offs[k + 1]], km["out_maps"][offs[k]:offs[k + 1]]):
                got.add((k, int(i), int(o)))
        tag = f"draw {it}: ks={ks} stride={stride} dil={dil} n={len(bc)}"
        assert got == pairs, tag
        # CSR lists ascend by output row inside every offset; the dense table agrees with them
        pt = km["pair_table"]
        assert int((pt >= 0).sum()) == len(pairs), tag
        for k in range(len(offs) - 1):
            assert np.all(np.diff(km["out_maps"][offs[k]:offs[k + 1]]) > 0), tag
        # invariant of the reference's tests: in = stride * out + offset[k]
        ko = km["kernel_offsets"]
        kidx = np.repeat(np.arange(len(offs) - 1), np.diff(offs))
        lhs = bc[km["in_maps"]][:, 1:]
        rhs = out_bc[km["out_maps"]][:, 1:] * np.asarray(stride) + ko[kidx]
        assert np.array_equal(lhs, rhs), tag
        assert np.array_equal(bc[km["in_maps"]][:, 0], out_bc[km["out_maps"]][:, 0]), tag


def test_reverse_table_inverts_the_pair_table():
    for it, ks, stride, dil, bc in _draws(12, 11):
        out_bc = bc if stride == (1, 1, 1) else okm.stride_coords(bc, stride)[0]
        pt = okm.generate_kernel_map(bc, out_bc, stride, ks, dilation=dil)["pair_table"]
        rev = okm.reverse_pair_table(pt, len(bc))
        k, o = np.nonzero(pt >= 0)
        assert np.array_equal(rev[k, pt[k, o]], o)
        assert int((rev >= 0).sum()) == len(k)


def test_stride_coords_is_the_sorted_set_of_floor_divided_cells():
    for it, ks, stride, dil, bc in _draws(12, 13):
        out, offsets = okm.stride_coords(bc, stride)
        cells = {(int(c[0]), *(int(v) // s if v >= 0 else -((-int(v) + s - 1) // s)
                               for v, s in zip(c[1:], stride))) for c in bc}
        assert [tuple(int(v) for v in r) for r in out] == sorted(cells)
        assert offsets[-1] == len(out) and np.all(np.diff(offsets) >= 0)


def _compress(masks, K):
    """The key compression of narrow_keys_iota_kernel (cuhash.cu) restated for K <= 32."""
    k = masks.astype(np.uint32)
    bits = K
    if K > 24:
        if K & 1:
            c = np.uint32(K // 2)
            k = ((k >> (c + np.uint32(1))) << c) | (k & ((np.uint32(1) << c) - np.uint32(1)))
            bits -= 1
        r = np.uint32(bits - 24)
        if r:
            k = (k >> r) ^ (k & ((np.uint32(1) << r) - np.uint32(1)))
    return k


@pytest.mark.parametrize("extent,seed", [(160, 0), (220, 3)])
def test_compressed_mask_keys_keep_the_plan_quality(extent, seed):
    bc = okm.batch_indexed([surface_coords(extent, seed)])
    pt = okm.generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3))["pair_table"]
    n = pt.shape[1]
    masks = np.zeros(n, np.uint32)
    for k in range(27):
        masks |= (pt[k] >= 0).astype(np.uint32) << np.uint32(k)
    keys = _compress(masks, 27)
    assert int(keys.max()) < (1 << 24)
    # equal masks share a key (rows of one mask stay adjacent); on a submanifold map the centre
    # bit is set everywhere, so nothing is lost by dropping it
    assert np.all(masks >> np.uint32(13) & np.uint32(1))
    first = {}
    for m, k in zip(masks.tolist(), keys.tolist()):
        assert first.setdefault(m, k) == k

    def steps(order, tile=256):
        m = np.concatenate([masks[order], np.zeros((-n) % tile, np.uint32)]).reshape(-1, tile)
        return sum(bin(int(u)).count("1") for u in np.bitwise_or.reduce(m, axis=1))

    full = steps(np.argsort(masks, kind="stable"))
    folded = steps(np.argsort(keys, kind="stable"))
    unsorted = steps(np.arange(n))
    assert folded <= 1.02 * full and folded < 0.6 * unsorted, (full, folded, unsorted)


def test_probes_outside_the_key_range_miss_like_the_dictionary_pin():
    """Voxels at opposite ends of the 18-bit range are NOT neighbours: the vectorised oracle agrees
    with the Python-dict enumeration (the reference tests' pin); ``wrap=True`` reproduces the
    masking of the reference's CUDA kernel, which aliases them, and differs only there."""
    lim = 131071
    c = np.array([[lim, lim, lim], [lim - 1, lim, lim], [-131072, -131072, -131072],
                  [-131071, -131072, -131072], [0, 0, 0], [lim, -131072, 0]], np.int32)
    bc = okm.batch_indexed([c])
    km = okm.generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3))
    got = {(k, int(i), int(o)) for k in range(27)
           for i, o in zip(km["in_maps"][km["offsets"][k]:km["offsets"][k + 1]],
                           km["out_maps"][km["offsets"][k]:km["offsets"][k + 1]])}
    assert got == okm.brute_force_pairs(bc, bc, (1, 1, 1), (3, 3, 3))
    wrapped = okm.generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3), wrap=True)
    assert int(wrapped["offsets"][-1]) > int(km["offsets"][-1])
    inner = okm.batch_indexed([c[[1, 3, 4]] // 2])
    a = okm.generate_kernel_map(inner, inner, (1, 1, 1), (3, 3, 3))
    b = okm.generate_kernel_map(inner, inner, (1, 1, 1), (3, 3, 3), wrap=True)
    assert np.array_equal(a["pair_table"], b["pair_table"])
