# SPDX-License-Identifier: Apache-2.0
"""Torch-tensor level wrappers over the C-ABI (device memory and streams only — no math here).

Every function enqueues on ``torch.cuda.current_stream()`` and never synchronises. Tensors are
allocated by the caller side (here) and handed to the library as raw pointers, the same ownership
rule the reference uses (detail/mask_gemm.py:723,874,934).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import check, lib

DTYPE_CODE = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}
TILE_M = 128


def _stream() -> int:
    # raw cudaStream_t of torch's current stream (torch.cuda.current_stream() costs ~15 us of
    # Python per call; the ctypes wrappers call this for every launch)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


_SM_COUNT = {}


def _sm_count(dev) -> int:
    idx = torch.device(dev).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _pc(t: Optional[Tensor], numel: int, what: str, like: Optional[Tensor] = None,
        dtype: torch.dtype = torch.float32):
    """Raw pointer of a per-channel vector / small buffer the kernels read or write with a FIXED
    element type: the dtype, element count, contiguity and device are checked here because the
    library only sees ``void*`` (a 2-byte buffer behind a ``float*`` is an out-of-bounds write)."""
    if t is None:
        return None
    if t.dtype != dtype or not t.is_contiguous() or t.numel() < numel or not t.is_cuda or (
            like is not None and t.device != like.device):
        raise _lib.WcnError(
            f"{what}: expected a contiguous {dtype} CUDA buffer of >= {numel} elements"
            f"{'' if like is None else ' on ' + str(like.device)}, got {t.dtype} "
            f"{tuple(t.shape)} on {t.device} (contiguous={t.is_contiguous()})")
    return t.data_ptr()


def _require_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "warpconvnet_b200 runs on CUDA (sm_100a) only; got a CPU tensor. "
                "There is no CPU fallback.")


def dtype_code(dtype: torch.dtype) -> int:
    try:
        return DTYPE_CODE[dtype]
    except KeyError:
        raise _lib.WcnError(f"unsupported feature dtype {dtype}; use bf16, fp16 or fp32") from None


# ------------------------------------------------------------------------------------------------
# hash table
# ------------------------------------------------------------------------------------------------
def next_power_of_2(n: int) -> int:
    return 1 if n <= 1 else 1 << (int(n) - 1).bit_length()


def hash_prepare(keys: Tensor, values: Tensor) -> None:
    _require_cuda(keys, values)
    check(lib.wcn_hash_prepare(_p(keys), _p(values), keys.numel(), _stream()), "hash_prepare")


def hash_insert(keys: Tensor, values: Tensor, coords: Tensor, status: Tensor) -> None:
    _require_cuda(keys, values, coords, status)
    assert coords.dtype == torch.int32 and coords.is_contiguous() and coords.shape[1] == 4
    check(lib.wcn_hash_insert(_p(keys), _p(values), _p(coords), coords.shape[0], keys.numel(),
                              _p(status), _stream()), "hash_insert")


def hash_search(keys: Tensor, values: Tensor, queries: Tensor) -> Tensor:
    _require_cuda(keys, values, queries)
    assert queries.dtype == torch.int32 and queries.is_contiguous() and queries.shape[1] == 4
    res = torch.empty(queries.shape[0], dtype=torch.int32, device=queries.device)
    check(lib.wcn_hash_search(_p(keys), _p(values), _p(queries), _p(res), queries.shape[0],
                              keys.numel(), _stream()), "hash_search")
    return res


# ------------------------------------------------------------------------------------------------
# kernel map
# ------------------------------------------------------------------------------------------------
def kernel_map_search(keys: Tensor, values: Tensor, out_coords: Tensor, offsets3: Tensor,
                      stride: Tuple[int, int, int], want_counts: bool = True,
                      want_mask: bool = True):
    """Returns (pair_table[K,M], block_counts[K,nb] | None, mask_keys[M] int64 | None)."""
    _require_cuda(keys, values, out_coords, offsets3)
    M = out_coords.shape[0]
    K = offsets3.shape[0]
    dev = out_coords.device
    nb = lib.wcn_kernel_map_num_blocks(M)
    pair_table = torch.empty((K, M), dtype=torch.int32, device=dev)
    block_counts = torch.empty((K, nb), dtype=torch.int32, device=dev) if want_counts else None
    mask_keys = torch.empty(M, dtype=torch.int64, device=dev) if want_mask else None
    check(lib.wcn_kernel_map_search(_p(keys), _p(values), keys.numel(), _p(out_coords), M,
                                    _p(offsets3), K, int(stride[0]), int(stride[1]),
                                    int(stride[2]), _p(pair_table), _p(block_counts),
                                    _p(mask_keys), _stream()), "kernel_map_search")
    return pair_table, block_counts, mask_keys


def kernel_map_stats(pair_table: Tensor, want_mask: bool = True):
    """(block_counts[K, nb], mask_keys[M] | None) of a finished pair table."""
    K, M = pair_table.shape
    dev = pair_table.device
    nb = lib.wcn_kernel_map_num_blocks(M)
    block_counts = torch.empty((K, nb), dtype=torch.int32, device=dev)
    mask_keys = torch.empty(M, dtype=torch.int64, device=dev) if want_mask else None
    check(lib.wcn_kernel_map_stats(_p(pair_table), K, M, _p(block_counts), _p(mask_keys),
                                   _stream()), "kernel_map_stats")
    return block_counts, mask_keys


def kernel_map_search_symmetric(keys: Tensor, values: Tensor, coords: Tensor, offsets3: Tensor,
                                status: Tensor, with_stats: bool = True):
    """Submanifold kernel map (in == out coordinates, odd kernel, stride 1): half the probes, hits
    mirrored. Returns (pair_table[K,M], block_counts[K,nb], mask_keys[M]); ``with_stats=False``
    returns the table only (None, None) — the caller runs kernel_map_stats where it needs it."""
    _require_cuda(keys, values, coords, offsets3, status)
    M, K = coords.shape[0], offsets3.shape[0]
    dev = coords.device
    nb = lib.wcn_kernel_map_num_blocks(M)
    pair_table = torch.empty((K, M), dtype=torch.int32, device=dev)
    check(lib.wcn_kernel_map_search_symmetric(_p(keys), _p(values), keys.numel(), _p(coords), M,
                                              _p(offsets3), K, _p(status), _p(pair_table),
                                              _stream()), "kernel_map_search_symmetric")
    if not with_stats:
        return pair_table, None, None
    block_counts, mask_keys = kernel_map_stats(pair_table)
    return pair_table, block_counts, mask_keys


def kernel_map_count(block_counts: Tensor) -> Tensor:
    """In-place block scan; returns device offsets[K+1] (int32)."""
    K, nb = block_counts.shape
    counts = torch.empty(K, dtype=torch.int32, device=block_counts.device)
    offsets = torch.empty(K + 1, dtype=torch.int32, device=block_counts.device)
    check(lib.wcn_kernel_map_count(_p(block_counts), K, nb, _p(counts), _p(offsets), _stream()),
          "kernel_map_count")
    return offsets


def kernel_map_scatter(pair_table: Tensor, block_prefix: Tensor, offsets_dev: Tensor,
                       num_pairs: int):
    K, M = pair_table.shape
    in_maps = torch.empty(num_pairs, dtype=torch.int32, device=pair_table.device)
    out_maps = torch.empty(num_pairs, dtype=torch.int32, device=pair_table.device)
    if num_pairs > 0:
        check(lib.wcn_kernel_map_scatter(_p(pair_table), _p(block_prefix), _p(offsets_dev),
                                         _p(in_maps), _p(out_maps), K, M, _stream()),
              "kernel_map_scatter")
    return in_maps, out_maps


def coords_unique(bcoords: Tensor, stride: Tuple[int, int, int], offsets3: Optional[Tensor],
                  n_batches: int, want_index: bool = False):
    """unique{(b, floor(xyz / stride) + offset_k)} over all rows and offsets, sorted by
    (b, x, y, z) — no host sync. Returns (rows [n * K, 4] int32 upper-bound buffer,
    first_index [n] int32 | None, meta [n_batches + 3] int32 = offsets[0..n_batches], total,
    status). Only the first ``total`` rows are valid."""
    _require_cuda(bcoords, offsets3)
    assert bcoords.dtype == torch.int32 and bcoords.is_contiguous() and bcoords.shape[1] == 4
    n = bcoords.shape[0]
    K = 1 if offsets3 is None else offsets3.shape[0]
    if offsets3 is not None:
        assert offsets3.dtype == torch.int32 and offsets3.is_contiguous() and offsets3.shape[1] == 3
    assert not (want_index and K != 1)
    dev = bcoords.device
    rows = torch.empty((max(n * K, 1), 4), dtype=torch.int32, device=dev)
    first = torch.empty(max(n, 1), dtype=torch.int32, device=dev) if want_index else None
    meta = torch.empty(n_batches + 3, dtype=torch.int32, device=dev)
    ws_bytes = lib.wcn_coords_unique_workspace_bytes(n * K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib.wcn_coords_unique(_p(bcoords), n, int(stride[0]), int(stride[1]), int(stride[2]),
                                _p(offsets3), K, n_batches, _p(rows), _p(first), _p(meta), _p(ws),
                                ws_bytes, _stream()), "coords_unique")
    return rows, first, meta


def reverse_pair_table(pair_table: Tensor, n_in: int) -> Tensor:
    K, M = pair_table.shape
    rev = torch.empty((K, n_in), dtype=torch.int32, device=pair_table.device)
    check(lib.wcn_reverse_pair_table(_p(pair_table), K, M, _p(rev), n_in, _stream()),
          "reverse_pair_table")
    return rev


def csr_to_pair_table(val_maps: Tensor, row_maps: Tensor, offsets_dev: Tensor, n_rows: int):
    K = offsets_dev.numel() - 1
    table = torch.empty((K, n_rows), dtype=torch.int32, device=val_maps.device)
    check(lib.wcn_csr_to_pair_table(_p(val_maps), _p(row_maps), _p(offsets_dev), K, n_rows,
                                    val_maps.numel(), _p(table), _stream()), "csr_to_pair_table")
    return table


def mask_keys(table: Tensor) -> Tensor:
    K, M = table.shape
    keys = torch.empty(M, dtype=torch.int64, device=table.device)
    check(lib.wcn_mask_keys(_p(table), K, M, _p(keys), _stream()), "mask_keys")
    return keys


@dataclass
class TilePlan:
    """Mask-sorted tiles of one [K, n_rows] neighbour table with their compact step lists
    (see wcn_build_tiles)."""
    step_nbr: Tensor   # [num_tiles, K, tile_rows] int32 (only the first tile_nk[t] steps are valid)
    step_k: Tensor     # [num_tiles, K] int32
    rows: Tensor       # [m_pad] int32
    tile_nk: Tensor    # [num_tiles] int32
    tile_cum: Tensor   # [num_tiles + 1] int32
    K: int
    n_rows: int
    m_pad: int
    num_tiles: int
    tile_rows: int
    cta_units: Optional[Tensor] = None  # [n_range_ctas + 1] unit range of every GEMM CTA
    n_range_ctas: int = 0


# 256-row tiles let two 128-row MMA sub-tiles share every weight slice (half the weight traffic);
# small inputs keep 128-row tiles so the tiles still spread over all SMs.
_TILE256_MIN_ROWS = 148 * 256


def build_tile_plan(table: Tensor, keys: Optional[Tensor] = None,
                    key_bits: Optional[int] = None, tile_rows: Optional[int] = None) -> TilePlan:
    """Replaces the reference's pair_mask + CUB argsort + per-tile mask OR
    (detail/mask_gemm.py:127-276)."""
    _require_cuda(table)
    K, M = table.shape
    dev = table.device
    if tile_rows is None:
        tile_rows = 256 if M >= _TILE256_MIN_ROWS else TILE_M
    rows_sorted = torch.empty(M, dtype=torch.int32, device=dev)
    ws_bytes = lib.wcn_sort_workspace_bytes(M)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    status = -2
    if keys is None and key_bits is None:
        # masks derived from the table inside the sort kernel (K <= 32, M <= 2^20)
        status = lib.wcn_sort_rows_by_table(_p(table), K, M, _p(rows_sorted), _p(ws), ws_bytes,
                                            _stream())
        if status not in (0, -2):
            check(status, "sort_rows_by_table")
    if status == -2:
        if keys is None:
            keys = mask_keys(table)
        check(lib.wcn_sort_rows_by_key(_p(keys), M, K if key_bits is None else key_bits,
                                       _p(rows_sorted), _p(ws), ws_bytes, _stream()),
              "sort_rows_by_key")
    m_pad = (M + tile_rows - 1) // tile_rows * tile_rows
    num_tiles = m_pad // tile_rows
    step_nbr = torch.empty((max(num_tiles, 1), K, tile_rows), dtype=torch.int32, device=dev)
    step_k = torch.empty((max(num_tiles, 1), K), dtype=torch.int32, device=dev)
    rows = torch.empty(max(m_pad, 1), dtype=torch.int32, device=dev)
    tile_nk = torch.empty(max(num_tiles, 1), dtype=torch.int32, device=dev)
    tile_cum = torch.empty(num_tiles + 1, dtype=torch.int32, device=dev)
    # per-CTA unit ranges of the gather-GEMM kernel for its default grid (one CTA per SM, never
    # more than 128-row units): computed by the plan's scan kernel, read by the GEMM prologue
    n_range = min(_sm_count(dev), num_tiles * (tile_rows // TILE_M))
    cta_units = torch.empty(n_range + 1, dtype=torch.int32, device=dev) if n_range > 0 else None
    # the 64-bit row masks (when the caller has them: kernel_map_stats / mask_keys) let every row
    # read only the table entries it has
    row_masks = keys if (keys is not None and key_bits is None and K <= 64
                         and keys.dtype in (torch.int64, torch.uint64)) else None
    check(lib.wcn_build_tiles_masked(_p(table), K, M, _p(rows_sorted), tile_rows, m_pad,
                                     _p(step_nbr), _p(step_k), _p(rows), _p(tile_nk),
                                     _p(tile_cum), n_range, _p(cta_units), _p(row_masks),
                                     _stream()),
          "build_tiles")
    return TilePlan(step_nbr, step_k, rows, tile_nk, tile_cum, K, M, m_pad, num_tiles, tile_rows,
                    cta_units, n_range)


# ------------------------------------------------------------------------------------------------
# GEMMs
# ------------------------------------------------------------------------------------------------
def weight_image(weight: Tensor, K: int, groups: int, cin_g: int, cout_g: int,
                 transpose_w: bool) -> Tensor:
    """weight: contiguous [K, groups, cin_g, cout_g] (or [K, cin, cout] when groups == 1)."""
    _require_cuda(weight)
    assert weight.is_contiguous()
    code = dtype_code(weight.dtype)
    nbytes = lib.wcn_weight_image_bytes(K, groups, cin_g, cout_g, code, int(transpose_w), None,
                                        None)
    if nbytes == 0:
        raise _lib.WcnError(
            f"unsupported channel configuration groups={groups} cin/g={cin_g} cout/g={cout_g}")
    img = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    check(lib.wcn_weight_image(_p(weight), _p(img), K, groups, cin_g, cout_g, code,
                               int(transpose_w), _stream()), "weight_image")
    return img


def weight_image_pair(weight: Tensor, K: int, groups: int, cin_g: int, cout_g: int,
                      image_dtype: torch.dtype, want_transposed: bool = True):
    """(forward image, dgrad image | None) of contiguous weights [K, groups, cin_g, cout_g] in ONE
    launch; fp32 weights are converted to a 16-bit ``image_dtype`` inside the kernel."""
    _require_cuda(weight)
    assert weight.is_contiguous()
    code = dtype_code(image_dtype)
    src = dtype_code(weight.dtype)
    nb_f = lib.wcn_weight_image_bytes(K, groups, cin_g, cout_g, code, 0, None, None)
    nb_t = lib.wcn_weight_image_bytes(K, groups, cin_g, cout_g, code, 1, None, None)
    if nb_f == 0 or (want_transposed and nb_t == 0):
        raise _lib.WcnError(
            f"unsupported channel configuration groups={groups} cin/g={cin_g} cout/g={cout_g}")
    img = torch.empty(nb_f, dtype=torch.uint8, device=weight.device)
    img_t = torch.empty(nb_t, dtype=torch.uint8, device=weight.device) if want_transposed else None
    check(lib.wcn_weight_image_pair(_p(weight), src, _p(img), _p(img_t), K, groups, cin_g, cout_g,
                                    code, _stream()), "weight_image_pair")
    return img, img_t


def gather_gemm(feats: Tensor, wimg: Tensor, plan: TilePlan, groups: int, cin_g: int,
                cout_g: int, out: Optional[Tensor] = None, bias: Optional[Tensor] = None,
                relu: bool = False, kflip: bool = False, max_ctas: int = 0,
                stats: Optional[Tensor] = None) -> Tensor:
    """out[plan.rows] = sum_k feats[plan.nbr[k]] @ W_k (forward AB / dgrad ABt gather-scatter).

    Every row of ``out`` (n_rows = plan.n_rows) is written exactly once, so ``out`` may be
    uninitialised memory (the reference zero-fills, detail/mask_gemm.py:723).
    ``stats``: zero-filled fp64 [2, groups * cout_g]; receives the per-channel sum / sum of squares
    of the stored output (the BatchNorm statistics, accumulated in the epilogue)."""
    _require_cuda(feats, wimg)
    assert feats.dim() == 2 and feats.stride(1) == 1
    code = dtype_code(feats.dtype)
    if out is None:
        out = torch.empty((plan.n_rows, groups * cout_g), dtype=feats.dtype, device=feats.device)
    assert out.stride(1) == 1 and out.dtype == feats.dtype
    _pc(bias, groups * cout_g, "bias", feats)
    check(lib.wcn_gather_gemm(_p(feats), feats.shape[0], feats.stride(0), _p(wimg), _p(out), out.stride(0),
                              _p(plan.step_nbr), _p(plan.step_k), _p(plan.rows), _p(plan.tile_nk),
                              _p(plan.tile_cum), plan.num_tiles, plan.tile_rows, plan.m_pad,
                              plan.K, groups, cin_g, cout_g, code, _p(bias), int(relu),
                              int(kflip), max_ctas, _p(plan.cta_units), plan.n_range_ctas,
                              _pc(stats, 2 * groups * cout_g, "stats", feats, torch.float64),
                              _stream()),
          "gather_gemm")
    return out


def wgrad(feats: Tensor, gout: Tensor, in_maps: Tensor, out_maps: Tensor, offsets_dev: Tensor,
          K: int, groups: int, cin_g: int, cout_g: int, dw: Optional[Tensor] = None,
          alpha: float = 1.0, unit_pairs: int = 0, max_ctas: int = 0,
          row_block_prefix: Optional[Tensor] = None, row_parts: int = 1,
          rounds: int = 1, identity_k: int = -1, status: Optional[Tensor] = None) -> Tensor:
    """dW[K, groups, cin_g, cout_g] (fp32) += X[in_maps]^T @ dY[out_maps] per offset.

    ``row_block_prefix`` ([K, n_row_blocks] int32, the scanned block counts of the kernel map)
    switches on the row-block-major unit order; ``identity_k`` (+ the hash table's ``status``
    word) marks the identity offset of a submanifold map, fetched as TMA tiles (see wcn_wgrad)."""
    _require_cuda(feats, gout, in_maps, out_maps, offsets_dev)
    assert feats.stride(1) == 1 and gout.stride(1) == 1 and feats.dtype == gout.dtype
    code = dtype_code(feats.dtype)
    if dw is None:
        dw = torch.zeros((K, groups, cin_g, cout_g), dtype=torch.float32, device=feats.device)
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    check(lib.wcn_wgrad(_p(feats), feats.stride(0), _p(gout), gout.stride(0), _p(dw), _p(in_maps),
                        _p(out_maps), _p(offsets_dev), K, groups, cin_g, cout_g, code,
                        ctypes.c_float(alpha), unit_pairs, max_ctas, _p(row_block_prefix),
                        0 if row_block_prefix is None else row_block_prefix.shape[1],
                        row_parts, rounds, identity_k, _p(status), feats.shape[0], gout.shape[0],
                        _stream()), "wgrad")
    return dw


# ------------------------------------------------------------------------------------------------
# k-nearest neighbours
# ------------------------------------------------------------------------------------------------
def knn_search(ref: Tensor, ref_offsets: Tensor, query: Tensor, query_offsets: Tensor, k: int,
               return_distances: bool = False):
    """int64 [M, k] global reference rows of the k nearest neighbours of every query point,
    searched inside the query's own batch item (ascending distance)."""
    _require_cuda(ref, query)
    assert ref.dtype == torch.float32 and query.dtype == torch.float32
    ref = ref.contiguous()
    query = query.contiguous()
    nb = ref_offsets.numel() - 1
    ro = ref_offsets.to(device=ref.device, dtype=torch.int32)
    qo = query_offsets.to(device=ref.device, dtype=torch.int32)
    out = torch.empty((query.shape[0], k), dtype=torch.int64, device=ref.device)
    dist = torch.empty((query.shape[0], k), dtype=torch.float32, device=ref.device) \
        if return_distances else None
    ws_bytes = lib.wcn_knn_workspace_bytes(ref.shape[0], nb)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=ref.device)
    check(lib.wcn_knn_search(_p(ref), ref.shape[0], _p(ro), _p(query), query.shape[0], _p(qo), nb,
                             k, _p(out), _p(dist), _p(ws), ws_bytes, _stream()), "knn_search")
    return (out, dist) if return_distances else out


def radius_search(ref: Tensor, ref_offsets: Tensor, query: Tensor, query_offsets: Tensor,
                  radius: float, return_distances: bool = True):
    """CSR neighbour lists of all reference points within ``radius`` of every query (same batch
    item): (int32 indices [Q] of GLOBAL reference rows, float32 distances [Q] | None, int64
    row_splits [M + 1]). One host sync (the total count sizes the outputs), like the reference."""
    _require_cuda(ref, query)
    assert ref.dtype == torch.float32 and query.dtype == torch.float32
    ref = ref.contiguous()
    query = query.contiguous()
    nb = ref_offsets.numel() - 1
    m = query.shape[0]
    ro = ref_offsets.to(device=ref.device, dtype=torch.int32)
    qo = query_offsets.to(device=ref.device, dtype=torch.int32)
    ws_bytes = lib.wcn_knn_workspace_bytes(ref.shape[0], nb)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=ref.device)
    counts = torch.empty(m, dtype=torch.int32, device=ref.device)
    check(lib.wcn_radius_count(_p(ref), ref.shape[0], _p(ro), _p(query), m, _p(qo), nb,
                               ctypes.c_float(radius), _p(counts), _p(ws), ws_bytes, _stream()),
          "radius_count")
    splits = torch.zeros(m + 1, dtype=torch.int64, device=ref.device)
    torch.cumsum(counts, dim=0, out=splits[1:])
    total = int(splits[-1].item()) if m > 0 else 0
    idx = torch.empty(total, dtype=torch.int32, device=ref.device)
    dist = torch.empty(total, dtype=torch.float32, device=ref.device) if return_distances else None
    if total > 0:
        check(lib.wcn_radius_fill(ref.shape[0], _p(query), m, _p(qo), nb, ctypes.c_float(radius),
                                  _p(splits), _p(idx), _p(dist), _p(ws), ws_bytes, _stream()),
              "radius_fill")
    return idx, dist, splits


# ------------------------------------------------------------------------------------------------
# per-channel normalisation / activation passes over the feature matrix (rownorm.cu)
# ------------------------------------------------------------------------------------------------
def _rows(t: Tensor):
    assert t.dim() == 2 and t.stride(1) == 1, "feature matrices are row-major [n, c]"
    return _p(t), t.stride(0)


def bn_stats(x: Tensor) -> Tensor:
    """fp64 [2, c]: per-channel sum and sum of squares of x [n, c]."""
    _require_cuda(x)
    n, c = x.shape
    sums = torch.zeros((2, c), dtype=torch.float64, device=x.device)
    px, ldx = _rows(x)
    check(lib.wcn_bn_stats(px, ldx, n, c, dtype_code(x.dtype), _p(sums), _stream()), "bn_stats")
    return sums


def bn_finalize(sums: Tensor, n: int, gamma: Optional[Tensor], beta: Optional[Tensor], eps: float,
                momentum: float, running_mean: Optional[Tensor], running_var: Optional[Tensor]):
    """(scale, shift, mean_rstd[2, c]) from the sums; updates the running statistics in place."""
    c = sums.shape[1]
    buf = torch.empty((4, c), dtype=torch.float32, device=sums.device)
    scale, shift, mean_rstd = buf[0], buf[1], buf[2:]
    check(lib.wcn_bn_finalize(_pc(sums, 2 * c, "sums", dtype=torch.float64), n, c,
                              _pc(gamma, c, "gamma", sums), _pc(beta, c, "beta", sums),
                              ctypes.c_float(eps), ctypes.c_float(momentum),
                              _pc(running_mean, c, "running_mean", sums),
                              _pc(running_var, c, "running_var", sums),
                              _p(scale), _p(shift), _p(mean_rstd), _stream()), "bn_finalize")
    return scale, shift, mean_rstd


def scale_shift_act(x: Tensor, scale: Tensor, shift: Tensor, residual: Optional[Tensor] = None,
                    relu: bool = False, out: Optional[Tensor] = None) -> Tensor:
    """act(x * scale[c] + shift[c] (+ residual))."""
    _require_cuda(x, scale, shift, residual)
    n, c = x.shape
    if out is None:
        out = torch.empty((n, c), dtype=x.dtype, device=x.device)
    px, ldx = _rows(x)
    po, ldo = _rows(out)
    pr, ldr = _rows(residual) if residual is not None else (None, 0)
    if residual is not None:
        assert residual.shape == x.shape and residual.dtype == x.dtype
    check(lib.wcn_scale_shift_act(px, ldx, pr, ldr, po, ldo, n, c, dtype_code(x.dtype),
                                  _pc(scale, c, "scale", x), _pc(shift, c, "shift", x),
                                  int(relu), _stream()), "scale_shift_act")
    return out


def bn_bwd_reduce(dy: Tensor, x: Tensor, y: Optional[Tensor], mean_rstd: Tensor,
                  mask_scale: Optional[Tensor] = None,
                  mask_shift: Optional[Tensor] = None) -> Tensor:
    """fp64 [2, c]: sum dz and sum dz * xhat with dz = dy * (y > 0) (y None: dz = dy);
    mask_scale / mask_shift: recompute the mask from x instead of reading y."""
    n, c = x.shape
    sums = torch.zeros((2, c), dtype=torch.float64, device=x.device)
    pd, ldd = _rows(dy)
    px, ldx = _rows(x)
    py, ldy = _rows(y) if y is not None else (None, 0)
    check(lib.wcn_bn_bwd_reduce(pd, ldd, px, ldx, py, ldy, n, c, dtype_code(x.dtype),
                                _pc(mean_rstd, 2 * c, "mean_rstd", x),
                                _pc(mask_scale, c, "mask_scale", x),
                                _pc(mask_shift, c, "mask_shift", x), _p(sums),
                                _stream()), "bn_bwd_reduce")
    return sums


def bn_bwd_apply(dy: Tensor, x: Optional[Tensor], y: Optional[Tensor], gamma: Tensor,
                 mean_rstd: Optional[Tensor], sums: Optional[Tensor], training: bool,
                 want_dres: bool, mask_scale: Optional[Tensor] = None,
                 mask_shift: Optional[Tensor] = None):
    """(dx, dres | None); see wcn_bn_bwd_apply."""
    n, c = dy.shape
    dx = torch.empty((n, c), dtype=dy.dtype, device=dy.device)
    dres = torch.empty((n, c), dtype=dy.dtype, device=dy.device) if want_dres else None
    pd, ldd = _rows(dy)
    px, ldx = _rows(x) if x is not None else (None, 0)
    py, ldy = _rows(y) if y is not None else (None, 0)
    pdx, lddx = _rows(dx)
    pdr, lddr = _rows(dres) if dres is not None else (None, 0)
    check(lib.wcn_bn_bwd_apply(pd, ldd, px, ldx, py, ldy, pdx, lddx, pdr, lddr, n, c,
                               dtype_code(dy.dtype), _pc(gamma, c, "gamma / scale", dy),
                               _pc(mean_rstd, 2 * c, "mean_rstd", dy),
                               _pc(sums, 2 * c, "sums", dy, torch.float64),
                               _pc(mask_scale, c, "mask_scale", dy),
                               _pc(mask_shift, c, "mask_shift", dy), int(training), _stream()),
          "bn_bwd_apply")
    return dx, dres


# ------------------------------------------------------------------------------------------------
# depthwise sparse convolution (conv_depthwise.cu)
# ------------------------------------------------------------------------------------------------
def depthwise_conv(feats: Tensor, weight: Tensor, table: Tensor, bias: Optional[Tensor] = None,
                   kflip: bool = False, relu: bool = False) -> Tensor:
    """out[r] = bias + sum_k feats[table[k, r]] * weight[k]; weight fp32 [K, C], table [K, M]."""
    _require_cuda(feats, weight, table, bias)
    K, M = table.shape
    C = weight.shape[1]
    assert weight.dtype == torch.float32 and weight.is_contiguous() and weight.shape[0] == K
    assert feats.shape[1] == C and table.dtype == torch.int32 and table.is_contiguous()
    out = torch.empty((M, C), dtype=feats.dtype, device=feats.device)
    pf, ldf = _rows(feats)
    po, ldo = _rows(out)
    check(lib.wcn_depthwise_conv(pf, ldf, po, ldo, _p(weight), _p(bias), _p(table), M, K, C,
                                 dtype_code(feats.dtype), int(kflip), int(relu), _stream()),
          "depthwise_conv")
    return out


def depthwise_conv_plan(feats: Tensor, weight: Tensor, plan: "TilePlan",
                        bias: Optional[Tensor] = None, kflip: bool = False,
                        relu: bool = False) -> Tensor:
    """depthwise_conv on the mask-sorted tile plan (compact step lists) instead of the dense
    [K, M] table; every row of the plan is written once."""
    _require_cuda(feats, weight, bias)
    K, C = weight.shape
    assert weight.dtype == torch.float32 and weight.is_contiguous() and K == plan.K
    assert feats.shape[1] == C
    out = torch.empty((plan.n_rows, C), dtype=feats.dtype, device=feats.device)
    pf, ldf = _rows(feats)
    po, ldo = _rows(out)
    check(lib.wcn_depthwise_conv_plan(pf, ldf, po, ldo, _p(weight), _p(bias), _p(plan.step_nbr),
                                      _p(plan.step_k), _p(plan.rows), _p(plan.tile_nk),
                                      plan.num_tiles, plan.tile_rows, K, C,
                                      dtype_code(feats.dtype), int(kflip), int(relu), _stream()),
          "depthwise_conv_plan")
    return out


def depthwise_wgrad(feats: Tensor, gout: Tensor, table: Tensor) -> Tensor:
    """fp32 dW[K, C] = sum_r feats[table[k, r]] * gout[r]."""
    _require_cuda(feats, gout, table)
    K, M = table.shape
    C = feats.shape[1]
    assert gout.shape == (M, C) and gout.dtype == feats.dtype
    dw = torch.zeros((K, C), dtype=torch.float32, device=feats.device)
    pf, ldf = _rows(feats)
    pg, ldg = _rows(gout)
    check(lib.wcn_depthwise_wgrad(pf, ldf, pg, ldg, _p(dw), _p(table), M, K, C,
                                  dtype_code(feats.dtype), _stream()), "depthwise_wgrad")
    return dw


def depthwise_wgrad_plan(feats: Tensor, gout: Tensor, plan: "TilePlan") -> Tensor:
    """fp32 dW[K, C] on the tile plan (gout rows are indexed through plan.rows)."""
    _require_cuda(feats, gout)
    C = feats.shape[1]
    assert gout.shape[1] == C and gout.dtype == feats.dtype
    dw = torch.zeros((plan.K, C), dtype=torch.float32, device=feats.device)
    pf, ldf = _rows(feats)
    pg, ldg = _rows(gout)
    check(lib.wcn_depthwise_wgrad_plan(pf, ldf, pg, ldg, _p(dw), _p(plan.step_nbr), _p(plan.step_k),
                                       _p(plan.rows), _p(plan.tile_nk), plan.num_tiles,
                                       plan.tile_rows, plan.K, C, dtype_code(feats.dtype),
                                       _stream()), "depthwise_wgrad_plan")
    return dw


def bn_forward(x: Tensor, gamma: Optional[Tensor], beta: Optional[Tensor], eps: float,
               momentum: float, running_mean: Optional[Tensor], running_var: Optional[Tensor],
               residual: Optional[Tensor], relu: bool, sums: Optional[Tensor] = None):
    """Training-mode BatchNorm (+ residual) (+ ReLU) in one native call.
    Returns (y, scale, shift, mean_rstd[2, c]). ``sums`` (fp64 [2, c]): per-channel sum / sum of
    squares of x already accumulated by the producing conv's epilogue — the statistics pass over
    x is skipped."""
    _require_cuda(x, residual)
    n, c = x.shape
    if sums is not None:
        scale, shift, mean_rstd = bn_finalize(sums, n, gamma, beta, eps, momentum, running_mean,
                                              running_var)
        return scale_shift_act(x, scale, shift, residual, relu), scale, shift, mean_rstd
    y = torch.empty((n, c), dtype=x.dtype, device=x.device)
    sums = torch.empty((2, c), dtype=torch.float64, device=x.device)
    buf = torch.empty((4, c), dtype=torch.float32, device=x.device)
    px, ldx = _rows(x)
    py, ldy = _rows(y)
    pr, ldr = _rows(residual) if residual is not None else (None, 0)
    if residual is not None:
        assert residual.shape == x.shape and residual.dtype == x.dtype
    check(lib.wcn_bn_forward(px, ldx, pr, ldr, py, ldy, n, c, dtype_code(x.dtype),
                             _pc(gamma, c, "gamma", x), _pc(beta, c, "beta", x),
                             ctypes.c_float(eps), ctypes.c_float(momentum),
                             _pc(running_mean, c, "running_mean", x),
                             _pc(running_var, c, "running_var", x), _p(sums), _p(buf), int(relu),
                             _stream()), "bn_forward")
    return y, buf[0], buf[1], buf[2:]


def bn_backward(dy: Tensor, x: Tensor, y: Optional[Tensor], gamma: Tensor, mean_rstd: Tensor,
                mask_scale: Optional[Tensor], mask_shift: Optional[Tensor], want_dres: bool):
    """(dx, dres | None, sums fp64 [2, c]: d beta, d gamma) of training-mode BatchNorm in one call."""
    n, c = x.shape
    dx = torch.empty((n, c), dtype=dy.dtype, device=dy.device)
    dres = torch.empty((n, c), dtype=dy.dtype, device=dy.device) if want_dres else None
    sums = torch.empty((2, c), dtype=torch.float64, device=x.device)
    pd, ldd = _rows(dy)
    px, ldx = _rows(x)
    py, ldy = _rows(y) if y is not None else (None, 0)
    pdx, lddx = _rows(dx)
    pdr, lddr = _rows(dres) if dres is not None else (None, 0)
    assert dy.shape == x.shape and dy.dtype == x.dtype
    check(lib.wcn_bn_backward(pd, ldd, px, ldx, py, ldy, pdx, lddx, pdr, lddr, n, c,
                              dtype_code(dy.dtype), _pc(gamma, c, "gamma", x),
                              _pc(mean_rstd, 2 * c, "mean_rstd", x),
                              _pc(mask_scale, c, "mask_scale", x),
                              _pc(mask_shift, c, "mask_shift", x), _p(sums), _stream()),
          "bn_backward")
    return dx, dres, sums
