# SPDX-License-Identifier: Apache-2.0
"""CPU oracle for the three sparse-conv GEMMs.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module; the product path (``warpconvnet_b200``) never does.

Restates the reference's explicit gather-matmul-scatter
(``warpconvnet/nn/functional/sparse_conv/detail/explicit.py``):

* forward  ``Y[out] += X[in] @ W_k``                      explicit.py:22-57
* dgrad    ``dX[in] += dY[out] @ W_k^T``                  explicit.py:60-101
* wgrad    ``dW_k  += X[in]^T @ dY[out]``                 explicit.py:95-97
* group conv = independent per-group explicit convs       tests/nn/test_sparse_conv.py:760-770
* 1x1x1 stride-1 shortcut ``feats @ W[0]``                helper.py:206-213

All math is done in float64 (or float32 on request) with torch CPU tensors: the oracle is the
high-precision value both the reference's kernels and ours are compared against
(tests/nn/test_kernel_correctness.py:64-65 uses fp32/fp64 explicit as truth).

Pinning: ``tests/golden/c1_conv.npz`` holds outputs of the REFERENCE's own
``_explicit_gemm_forward_logic`` / ``_explicit_gemm_backward_logic`` (imported from
/root/reference with two stub modules by ``tests/golden/make_golden.py``); ``tests/test_oracle.py``
checks this oracle against them, and the ones-KAT of
``scripts/validate_tiles_on_device.py:46-96`` (x = 1, w = 1 => Y[r,:] = Cin * degree(r)).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch


def _t(a, dtype):
    if isinstance(a, torch.Tensor):
        return a.detach().to("cpu").to(dtype)
    return torch.as_tensor(np.asarray(a)).to(dtype)


def forward(x, w, in_maps, out_maps, offsets, n_out: int, dtype=torch.float64) -> torch.Tensor:
    """Y[n_out, Cout].  w: [K, Cin, Cout]."""
    x = _t(x, dtype)
    w = _t(w, dtype)
    im = torch.as_tensor(np.asarray(in_maps)).long()
    om = torch.as_tensor(np.asarray(out_maps)).long()
    offs = np.asarray(offsets)
    y = torch.zeros(n_out, w.shape[-1], dtype=dtype)
    for k in range(w.shape[0]):
        s, e = int(offs[k]), int(offs[k + 1])
        if e > s:
            y.index_add_(0, om[s:e], x[im[s:e]] @ w[k])
    return y


def backward(gy, x, w, in_maps, out_maps, offsets, dtype=torch.float64
             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(dX[n_in, Cin], dW[K, Cin, Cout])."""
    gy = _t(gy, dtype)
    x = _t(x, dtype)
    w = _t(w, dtype)
    im = torch.as_tensor(np.asarray(in_maps)).long()
    om = torch.as_tensor(np.asarray(out_maps)).long()
    offs = np.asarray(offsets)
    dx = torch.zeros_like(x)
    dw = torch.zeros_like(w)
    for k in range(w.shape[0]):
        s, e = int(offs[k]), int(offs[k + 1])
        if e > s:
            g = gy[om[s:e]]
            dx.index_add_(0, im[s:e], g @ w[k].T)
            dw[k] = x[im[s:e]].T @ g
    return dx, dw


def forward_grouped(x, w, in_maps, out_maps, offsets, n_out: int, dtype=torch.float64):
    """w: [K, G, Cin/G, Cout/G]; per-group explicit conv, outputs concatenated on channels."""
    w = _t(w, dtype)
    x = _t(x, dtype)
    K, G, cg, og = w.shape
    outs = []
    for g in range(G):
        outs.append(forward(x[:, g * cg:(g + 1) * cg], w[:, g], in_maps, out_maps, offsets, n_out,
                            dtype))
    return torch.cat(outs, dim=1)


def backward_grouped(gy, x, w, in_maps, out_maps, offsets, dtype=torch.float64):
    w = _t(w, dtype)
    x = _t(x, dtype)
    gy = _t(gy, dtype)
    K, G, cg, og = w.shape
    dxs, dws = [], []
    for g in range(G):
        dx, dw = backward(gy[:, g * og:(g + 1) * og], x[:, g * cg:(g + 1) * cg], w[:, g], in_maps,
                          out_maps, offsets, dtype)
        dxs.append(dx)
        dws.append(dw)
    return torch.cat(dxs, dim=1), torch.stack(dws, dim=1)


def degree(out_maps, n_out: int) -> np.ndarray:
    """Number of (offset, input) pairs feeding each output row (ones-KAT helper)."""
    return np.bincount(np.asarray(out_maps), minlength=n_out)


def rel_max_err(a, ref) -> float:
    """max|a-ref| / max|ref| — the reference's kernel-correctness metric
    (tests/nn/test_kernel_correctness.py:140-146)."""
    a = _t(a, torch.float64)
    ref = _t(ref, torch.float64)
    denom = float(ref.abs().max())
    return float((a - ref).abs().max()) / (denom if denom > 0 else 1.0)


def rdiff(a, ref) -> float:
    """mean-relative difference used by tests/nn/test_mask_gemm_numerical.py:39-40."""
    a = _t(a, torch.float64)
    ref = _t(ref, torch.float64)
    denom = float(ref.abs().mean())
    return float((a - ref).abs().mean()) / (denom if denom > 0 else 1.0)


# ------------------------------------------------------------------------------------------------
# depthwise sparse convolution (weight [K, C]) — restates the reference's explicit path
# warpconvnet/nn/functional/sparse_conv_depth.py:227-257 (forward), :260-308 (backward).
# Pinned by tests/golden/dw_*.npz = outputs of the reference's own
# _explicit_depthwise_forward_logic / _explicit_depthwise_backward_logic (make_golden.py).
# ------------------------------------------------------------------------------------------------
def depthwise_forward(x, w, in_maps, out_maps, offsets, n_out: int, dtype=torch.float64):
    """Y[n_out, C] = sum_k X[in_k] * w[k]."""
    x = _t(x, dtype)
    w = _t(w, dtype)
    im = torch.as_tensor(np.asarray(in_maps)).long()
    om = torch.as_tensor(np.asarray(out_maps)).long()
    offs = np.asarray(offsets)
    y = torch.zeros(n_out, w.shape[-1], dtype=dtype)
    for k in range(w.shape[0]):
        s, e = int(offs[k]), int(offs[k + 1])
        if e > s:
            y.index_add_(0, om[s:e], x[im[s:e]] * w[k].unsqueeze(0))
    return y


def depthwise_backward(gy, x, w, in_maps, out_maps, offsets, dtype=torch.float64):
    """(dX[n_in, C], dW[K, C])."""
    gy = _t(gy, dtype)
    x = _t(x, dtype)
    w = _t(w, dtype)
    im = torch.as_tensor(np.asarray(in_maps)).long()
    om = torch.as_tensor(np.asarray(out_maps)).long()
    offs = np.asarray(offsets)
    dx = torch.zeros_like(x)
    dw = torch.zeros_like(w)
    for k in range(w.shape[0]):
        s, e = int(offs[k]), int(offs[k + 1])
        if e > s:
            g = gy[om[s:e]]
            dx.index_add_(0, im[s:e], g * w[k].unsqueeze(0))
            dw[k] = (x[im[s:e]] * g).sum(dim=0)
    return dx, dw
