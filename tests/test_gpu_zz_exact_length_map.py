# SPDX-License-Identifier: Apache-2.0
"""The kernel-map builder's second allocation mode. Above ``_DEFERRED_MAX_PAIRS`` (K x M > 2^25,
e.g. the 2.4 M-voxel level of the MinkUNet-14 shape) the CSR lists are allocated at their exact
length after one host read of the offsets instead of into upper-bound buffers
(geometry/coords/search/torch_discrete.py). Forced here on a small input; the result must be the
same bit-exact map as the oracle's and drive the conv kernels like the deferred one."""
import numpy as np
import pytest
import torch

from conftest import random_coords
from oracle import conv as oconv
from oracle import kernel_map as okm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("stride,ks", [(1, 3), (2, 2)])
def test_exact_length_allocation_mode_matches_oracle(monkeypatch, stride, ks):
    from warpconvnet_b200.geometry.coords.search import torch_discrete as td
    from warpconvnet_b200.nn.functional.sparse_conv import sparse_conv_forward
    monkeypatch.setattr(td, "_DEFERRED_MAX_PAIRS", 0)
    bc = okm.batch_indexed([random_coords(6000, 0.3, 0), random_coords(2500, 0.2, 1)])
    out_bc = bc if stride == 1 else okm.stride_coords(bc, (stride,) * 3)[0]
    ti = torch.from_numpy(np.ascontiguousarray(bc)).cuda()
    to = ti if stride == 1 else torch.from_numpy(np.ascontiguousarray(out_bc)).cuda()
    km = td.generate_kernel_map(ti, to, (stride,) * 3, (ks,) * 3)
    ref = okm.generate_kernel_map(bc, out_bc, (stride,) * 3, (ks,) * 3)
    assert np.array_equal(km.offsets.numpy(), ref["offsets"])
    assert km.in_maps.shape[0] == int(ref["offsets"][-1])          # exact length, no slack
    assert np.array_equal(km.in_maps.cpu().numpy(), ref["in_maps"])
    assert np.array_equal(km.out_maps.cpu().numpy(), ref["out_maps"])
    assert np.array_equal(km._pair_table.cpu().numpy(), ref["pair_table"])
    g = torch.Generator().manual_seed(3)
    x = torch.randn(len(bc), 32, generator=g).cuda().bfloat16()
    w = (torch.randn(ks ** 3, 32, 64, generator=g) * (ks ** 3 * 32) ** -0.5).cuda().bfloat16()
    y = sparse_conv_forward(x, w, km, len(out_bc))
    y_ref = oconv.forward(x.float().cpu(), w.float().cpu(), ref["in_maps"], ref["out_maps"],
                          ref["offsets"], len(out_bc))
    assert oconv.rel_max_err(y, y_ref) < 1e-2
