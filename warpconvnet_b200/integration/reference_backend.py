# SPDX-License-Identifier: Apache-2.0
"""Seam A of INTEGRATION.md as running code: register this library as a backend of the
REFERENCE's own sparse-conv dispatcher.

The reference keeps ``FORWARD_BACKENDS: Dict[str, Callable[[FwdCtx], Tensor | int]]`` and
``BACKWARD_BACKENDS: Dict[str, Callable[[BwdCtx], Tuple[Tensor | int | None, Tensor | None]]]``
(warpconvnet/nn/functional/sparse_conv/detail/backends.py:90-131,443-463) and rejects algorithm
names that are in neither its adaptive pool nor ``_ALL_AB_PARAMS`` / ``_ALL_ATB_PARAMS``
(detail/algo_params.py:1104-1129). ``register()`` adds one adapter per direction plus the pool
entries, after which reference user code selects the kernels of this library with

    SparseConv3d(cin, cout, 3, fwd_algo=["wcn_b200"], dgrad_algo=["wcn_b200"],
                 wgrad_algo=["wcn_b200"])

(a list bypasses the enum parse, nn/modules/sparse_conv.py:130-137). The adapters follow the
reference's result convention: a tensor on success, a negative ``int`` status when the shape is
not supported ("try another candidate", backends.py:489-510). The reference's kernel map (CSR
``in_maps`` / ``out_maps`` on the device, ``offsets`` on the CPU, search_results.py:55-202) is
wrapped once per map into this library's ``IntSearchResult`` — which derives the pair tables and
tile plans through ``wcn_csr_to_pair_table`` / ``wcn_build_tiles`` — and cached on the map object
like the reference caches its own ``_mask_data``.

Nothing in this repository imports this module; it needs the reference package on ``sys.path``.
"""
from __future__ import annotations

from typing import Optional

import torch

from warpconvnet_b200._lib import WcnError
from warpconvnet_b200.geometry.coords.search.search_results import IntSearchResult
from warpconvnet_b200.nn.functional.sparse_conv import (sparse_conv_dgrad, sparse_conv_forward,
                                                        sparse_conv_wgrad)

BACKEND_NAME = "wcn_b200"
STATUS_UNSUPPORTED = -1  # "kErrorProblemNotSupported" slot of csrc/include/gemm_error_codes.h:7-15
_CACHE_ATTR = "_wcn_b200_map"


def wrap_kernel_map(kernel_map) -> IntSearchResult:
    """This library's view of a reference ``IntSearchResult`` (built once, cached on it)."""
    own = getattr(kernel_map, _CACHE_ATTR, None)
    if own is None:
        own = IntSearchResult(kernel_map.in_maps.int(), kernel_map.out_maps.int(),
                              kernel_map.offsets, kernel_map.identity_map_index)
        setattr(kernel_map, _CACHE_ATTR, own)
    return own


def _cast(t: torch.Tensor, dtype: Optional[torch.dtype]) -> torch.Tensor:
    return t if dtype is None or t.dtype == dtype else t.to(dtype)


def forward_adapter(ctx):
    """``FwdCtx -> Tensor | int`` (replaces ``_fwd_mask``, backends.py:204-212)."""
    x = _cast(ctx.in_features, ctx.compute_dtype)
    w = _cast(ctx.weight, ctx.compute_dtype)
    try:
        y = sparse_conv_forward(x, w, wrap_kernel_map(ctx.kernel_map), ctx.num_out_coords,
                                groups=ctx.groups)
    except (WcnError, ValueError):  # unsupported shape / channel configuration
        return STATUS_UNSUPPORTED
    return y.to(ctx.in_features.dtype)


def backward_adapter(ctx):
    """``BwdCtx -> (grad_in | int | None, grad_weight | None)`` (replaces ``_bwd_mask``,
    backends.py:429-431). ``needs_input_grad`` = (features, weight, ...)."""
    gy = _cast(ctx.grad_output, ctx.compute_dtype)
    x = _cast(ctx.in_features, ctx.compute_dtype)
    w = _cast(ctx.weight, ctx.compute_dtype)
    km = wrap_kernel_map(ctx.kernel_map)
    grad_in = grad_w = None
    try:
        if ctx.needs_input_grad[0]:
            grad_in = sparse_conv_dgrad(gy, w, km, ctx.in_features.shape[0],
                                        groups=ctx.groups).to(ctx.in_features.dtype)
        if len(ctx.needs_input_grad) > 1 and ctx.needs_input_grad[1]:
            grad_w = sparse_conv_wgrad(x, gy, tuple(ctx.weight.shape), km,
                                       groups=ctx.groups).to(ctx.weight.dtype)
    except (WcnError, ValueError):
        return STATUS_UNSUPPORTED, None
    return grad_in, grad_w


def register(name: str = BACKEND_NAME) -> str:
    """Adds the adapters to the reference's registries and the name to its parameter pools.
    Idempotent. Returns the registered name."""
    from warpconvnet.nn.functional.sparse_conv.detail import algo_params, backends
    backends.FORWARD_BACKENDS[name] = forward_adapter
    backends.BACKWARD_BACKENDS[name] = backward_adapter
    for pool in (algo_params._ALL_AB_PARAMS, algo_params._ALL_ATB_PARAMS):
        if not any(str(tag) == name for tag, _ in pool):
            pool.append((name, {}))
    return name
