# SPDX-License-Identifier: Apache-2.0
"""ctypes binding of ``csrc/libwcn_b200.so`` (the C-ABI declared in ``include/wcn_b200.h``).

There is NO fallback: if the shared library is missing the import fails loudly, and every wrapper
raises on a non-zero status (the reference raises ``RuntimeError`` on a negative GemmStatus,
``warpconvnet/nn/functional/sparse_conv/detail/backends.py:489-510``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libwcn_b200.so")

_ERRORS = {
    -1: "invalid argument",
    -2: "unsupported shape",
    -3: "misaligned pointer or stride",
    -4: "unsupported dtype",
    -5: "CUDA launch error",
    -6: "workspace too small",
}


class WcnError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"warpconvnet_b200: native library not found at {LIB_PATH}. Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` or "
            f"`warpconvnet_b200/csrc/build.sh`. There is no CPU / PyTorch fallback."
        )
    return ctypes.CDLL(LIB_PATH)


lib = _load()

# name -> (restype, argtypes); must list every symbol of include/wcn_b200.h
SIGNATURES = {
    "wcn_version": (c_char_p, []),
    "wcn_built_for_sm100a": (c_int, []),
    "wcn_launch_count": (c_longlong, []),
    "wcn_hash_prepare": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "wcn_hash_insert": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "wcn_hash_search": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "wcn_kernel_map_num_blocks": (c_int, [c_int]),
    "wcn_kernel_map_search": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                      c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "wcn_kernel_map_search_symmetric": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                                c_int, c_void_p, c_void_p, c_void_p]),
    "wcn_kernel_map_stats": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "wcn_kernel_map_count": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "wcn_kernel_map_scatter": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                       c_int, c_void_p]),
    "wcn_reverse_pair_table": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "wcn_csr_to_pair_table": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                      c_void_p]),
    "wcn_mask_keys": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "wcn_coords_unique_workspace_bytes": (c_size_t, [c_longlong]),
    "wcn_coords_unique": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "wcn_sort_workspace_bytes": (c_size_t, [c_int]),
    "wcn_sort_rows_by_key": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    "wcn_sort_rows_by_table": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                       c_void_p]),
    "wcn_build_tiles": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "wcn_build_tiles_masked": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                       c_void_p, c_void_p]),
    "wcn_knn_workspace_bytes": (c_size_t, [c_int, c_int]),
    "wcn_knn_search": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "wcn_radius_count": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_float,
                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "wcn_radius_fill": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_void_p]),
    "wcn_weight_image_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int),
                                          POINTER(c_int)]),
    "wcn_weight_image": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p]),
    "wcn_weight_image_pair": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_int, c_void_p]),
    "wcn_gather_gemm": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_void_p, c_longlong, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                c_void_p, c_int, c_void_p, c_void_p]),
    "wcn_bn_forward": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int,
                               c_int, c_int, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, c_void_p]),
    "wcn_bn_backward": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong,
                                c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "wcn_peer_allreduce_flag_words": (c_int, []),
    "wcn_peer_allreduce_timeout_word": (c_int, []),
    "wcn_peer_allreduce_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_longlong, c_float,
                                       c_int, c_void_p]),
    "wcn_depthwise_conv": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p,
                                   c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "wcn_depthwise_conv_plan": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                        c_int, c_int, c_int, c_int, c_void_p]),
    "wcn_depthwise_wgrad_plan": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_int, c_int, c_void_p]),
    "wcn_depthwise_wgrad": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p,
                                    c_int, c_int, c_int, c_int, c_void_p]),
    "wcn_bn_stats": (c_int, [c_void_p, c_longlong, c_int, c_int, c_int, c_void_p, c_void_p]),
    "wcn_bn_finalize": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_float,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "wcn_scale_shift_act": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong,
                                    c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "wcn_bn_bwd_reduce": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong,
                                  c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "wcn_bn_bwd_apply": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong,
                                 c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                 c_void_p]),
    "wcn_wgrad": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p,
                          c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int,
                          c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_longlong, c_longlong,
                          c_void_p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header / library drift
    _fn.restype = _res
    _fn.argtypes = _args


def check(status: int, what: str) -> None:
    if status != 0:
        raise WcnError(f"{what} failed: {_ERRORS.get(status, 'unknown error')} (status {status})")


def version() -> str:
    return lib.wcn_version().decode()
