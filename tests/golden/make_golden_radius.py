# SPDX-License-Identifier: Apache-2.0
"""Golden fixtures for the radius search: outputs of the REFERENCE's own ``radius_search``
(warpconvnet/geometry/coords/search/radius.py:162-225, its torch ``cdist`` path) per batch item on
CPU, assembled exactly like its ``batched_radius_search`` (:228-291).  TEST INFRASTRUCTURE ONLY;
runs only where /root/reference is mounted.   Usage:  python tests/golden/make_golden_radius.py
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import points as opts  # noqa: E402
from oracle import ref_adapter  # noqa: E402


def main():
    ref_adapter.load()
    rmod = importlib.import_module("warpconvnet.geometry.coords.search.radius")
    g = torch.Generator().manual_seed(3)
    for name, sizes, qsizes, r in (("radius_b2", (700, 500), (300, 400), 0.11),
                                   ("radius_b1_self", (900,), (900,), 0.08)):
        ref = torch.rand(sum(sizes), 3, generator=g)
        query = ref.clone() if name.endswith("self") else torch.rand(sum(qsizes), 3, generator=g)
        ro = np.concatenate([[0], np.cumsum(sizes)])
        qo = np.concatenate([[0], np.cumsum(qsizes)])
        idx_l, dist_l, split_l, off = [], [], [], 0
        for b in range(len(sizes)):
            i, d, s = rmod.radius_search(ref[ro[b]:ro[b + 1]], query[qo[b]:qo[b + 1]], r)
            idx_l.append(i.long() + int(ro[b]))
            dist_l.append(d)
            split_l.append((s if b == len(sizes) - 1 else s[:-1]).long() + off)
            off += len(i)
        idx, dist, splits = torch.cat(idx_l).numpy(), torch.cat(dist_l).numpy(), torch.cat(split_l).numpy()
        # the fixture must not depend on fp32 rounding at the radius: the fp64 oracle agrees exactly
        oi, od, osplit = opts.radius(ref.numpy(), ro, query.numpy(), qo, r)
        assert np.array_equal(osplit, splits)
        for q in range(len(splits) - 1):
            assert np.array_equal(np.sort(idx[splits[q]:splits[q + 1]]), oi[osplit[q]:osplit[q + 1]])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), ref=ref.numpy(), query=query.numpy(),
                            ref_offsets=ro, query_offsets=qo, radius=np.float32(r), idx=idx,
                            dist=dist, splits=splits)
        print("wrote", name, "pairs =", len(idx))


if __name__ == "__main__":
    main()
