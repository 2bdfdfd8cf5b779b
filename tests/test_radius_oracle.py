# SPDX-License-Identifier: Apache-2.0
"""CPU: the brute-force radius oracle (oracle/points.py) against the committed outputs of the
reference's own ``radius_search`` (tests/golden/radius_*.npz, make_golden_radius.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import points as opts

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "radius_*.npz")))


def test_fixtures_present():
    assert len(GOLDEN) == 2


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_radius_oracle_matches_reference_outputs(path):
    d = np.load(path)
    idx, dist, splits = opts.radius(d["ref"], d["ref_offsets"], d["query"], d["query_offsets"],
                                    float(d["radius"]))
    assert np.array_equal(splits, d["splits"])
    for q in range(len(splits) - 1):
        s, e = splits[q], splits[q + 1]
        order = np.argsort(d["idx"][s:e])
        assert np.array_equal(d["idx"][s:e][order], idx[s:e])
        # the reference computes fp32 cdist through a matmul: up to 7e-4 absolute error at zero distance
        assert np.allclose(d["dist"][s:e][order], dist[s:e], atol=1e-3)
