# SPDX-License-Identifier: Apache-2.0
"""CPU: the depthwise oracle (oracle/conv.py) against the committed outputs of the reference's own
explicit depthwise path (tests/golden/dw_*.npz, made by make_golden_depthwise.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import conv as oconv

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dw_*.npz")))


def test_fixtures_present():
    assert len(GOLDEN) == 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_depthwise_oracle_matches_reference_outputs(path):
    d = np.load(path)
    y = oconv.depthwise_forward(d["x"], d["w"], d["in_maps"], d["out_maps"], d["offsets"],
                                len(d["out_bc"]))
    dx, dw = oconv.depthwise_backward(d["gy"], d["x"], d["w"], d["in_maps"], d["out_maps"],
                                      d["offsets"])
    for ours, ref in ((y, d["y"]), (dx, d["dx"]), (dw, d["dw"])):
        ref = torch.from_numpy(ref)
        assert float((ours - ref).abs().max()) <= 1e-12 * max(1.0, float(ref.abs().max()))
