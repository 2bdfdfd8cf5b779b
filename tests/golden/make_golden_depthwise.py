# SPDX-License-Identifier: Apache-2.0
"""Golden fixtures for the depthwise conv: outputs of the REFERENCE's own
``_explicit_depthwise_forward_logic`` / ``_explicit_depthwise_backward_logic``
(warpconvnet/nn/functional/sparse_conv_depth.py:227-308) in fp64 on CPU, on the kernel maps of
``oracle.kernel_map``.  TEST INFRASTRUCTURE ONLY; runs only where /root/reference is mounted.

Usage:  python tests/golden/make_golden_depthwise.py
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import c1_coords, toy_coords  # noqa: E402
from oracle import kernel_map as okm  # noqa: E402
from oracle import ref_adapter  # noqa: E402


def main():
    _, _, ISR = ref_adapter.load()
    dwm = importlib.import_module("warpconvnet.nn.functional.sparse_conv_depth")
    cases = [("dw_c1_s1_k3", c1_coords(), (1, 1, 1), (3, 3, 3), 8),
             ("dw_c1_s2_k2", c1_coords(), (2, 2, 2), (2, 2, 2), 12),
             ("dw_toy_s1_k3", toy_coords(), (1, 1, 1), (3, 3, 3), 4)]
    for name, in_bc, stride, ks, c in cases:
        out_bc = in_bc if all(s == 1 for s in stride) else okm.stride_coords(in_bc, stride)[0]
        km = okm.generate_kernel_map(in_bc, out_bc, stride, ks)
        K = int(np.prod(ks))
        g = torch.Generator().manual_seed(len(name))
        x = torch.randn(len(in_bc), c, generator=g, dtype=torch.float64)
        w = torch.randn(K, c, generator=g, dtype=torch.float64) * K ** -0.5
        gy = torch.randn(len(out_bc), c, generator=g, dtype=torch.float64)
        ref_km = ISR(torch.from_numpy(km["in_maps"]).long(), torch.from_numpy(km["out_maps"]).long(),
                     torch.from_numpy(km["offsets"]).long(),
                     identity_map_index=km["identity_map_index"])
        y = dwm._explicit_depthwise_forward_logic(x, w, ref_km, len(out_bc))
        dx, dw = dwm._explicit_depthwise_backward_logic(gy, x, w, ref_km, device=torch.device("cpu"))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), in_bc=in_bc, out_bc=out_bc,
                            stride=np.array(stride), ksize=np.array(ks), x=x.numpy(), w=w.numpy(),
                            gy=gy.numpy(), in_maps=km["in_maps"], out_maps=km["out_maps"],
                            offsets=km["offsets"], y=y.numpy(), dx=dx.numpy(), dw=dw.numpy())
        print("wrote", name, "L =", int(km["offsets"][-1]))


if __name__ == "__main__":
    main()
