# SPDX-License-Identifier: Apache-2.0
"""Autograd function of the sparse convolution
(drop-in for ``UnifiedSpatiallySparseConvFunction``,
warpconvnet/nn/functional/sparse_conv/detail/unified.py:143-785).

There is exactly one backend — the hand-written sm_100a kernels behind ``libwcn_b200.so`` — so the
reference's per-call autotune / benchmark cache / fallback chain is gone by design. The algo-mode
enums are kept so reference user code keeps working; every value maps to the same kernels.

  forward : Y  = gather_gemm(X,  W image,      fwd tile plan)            AB_gather_scatter
  dgrad   : dX = gather_gemm(dY, W^T image,    bwd tile plan)            ABt_gather_scatter
  wgrad   : dW = wgrad(X, dY, CSR pair lists)  fp32, cast to W's dtype   AtB_gather_gather
"""
from __future__ import annotations

from enum import Enum
from typing import Optional


import torch
from torch import Tensor
from torch.autograd import Function

from warpconvnet_b200 import _ops
from warpconvnet_b200.geometry.coords.search.search_results import (IntSearchResult,
                                                                    check_pending_kernel_maps)


class SPARSE_CONV_AB_ALGO_MODE(Enum):
    EXPLICIT_GEMM = "explicit_gemm"
    IMPLICIT_GEMM = "implicit_gemm"
    CUTLASS_IMPLICIT_GEMM = "cutlass_implicit_gemm"
    CUTE_IMPLICIT_GEMM = "cute_implicit_gemm"
    EXPLICIT_GEMM_GROUPED = "explicit_gemm_grouped"
    IMPLICIT_GEMM_GROUPED = "implicit_gemm_grouped"
    CUTLASS_GROUPED_HYBRID = "cutlass_grouped_hybrid"
    CUTE_GROUPED = "cute_grouped"
    MASK_GEMM = "mask_gemm"
    TCGEN05 = "tcgen05"  # what actually runs
    AUTO = "auto"
    ALL = "all"
    TRIMMED = "trimmed"


class SPARSE_CONV_ATB_ALGO_MODE(Enum):
    EXPLICIT_GEMM = "explicit_gemm"
    IMPLICIT_GEMM = "implicit_gemm"
    CUTLASS_IMPLICIT_GEMM = "cutlass_implicit_gemm"
    CUTE_IMPLICIT_GEMM = "cute_implicit_gemm"
    EXPLICIT_GEMM_GROUPED = "explicit_gemm_grouped"
    IMPLICIT_GEMM_GROUPED = "implicit_gemm_grouped"
    CUTLASS_GROUPED_HYBRID = "cutlass_grouped_hybrid"
    CUTE_GROUPED = "cute_grouped"
    MASK_GEMM = "mask_gemm"
    TCGEN05 = "tcgen05"
    AUTO = "auto"
    ALL = "all"
    TRIMMED = "trimmed"


_CH_ALIGN = 16  # channels are padded to a multiple of 16 (one 32-byte MMA K-slice of bf16)


def _round_up(v: int, a: int) -> int:
    return (v + a - 1) // a * a


def _pad_cols(t: Tensor, cols: int) -> Tensor:
    if t.shape[1] == cols and t.stride(1) == 1 and (t.stride(0) * t.element_size()) % 16 == 0 \
            and t.data_ptr() % 16 == 0:
        return t
    out = torch.zeros((t.shape[0], cols), dtype=t.dtype, device=t.device)
    out[:, : t.shape[1]] = t
    return out


def _canon_weight(weight: Tensor, groups: int):
    """-> (w4 [K, G, cin_g, cout_g] contiguous, padded), cin_g, cout_g, real cin_g, real cout_g"""
    if groups == 1:
        K, cin, cout = weight.shape
        cin_p, cout_p = _round_up(cin, _CH_ALIGN), _round_up(cout, _CH_ALIGN)
        if cin_p != cin or cout_p != cout:
            w = torch.zeros((K, cin_p, cout_p), dtype=weight.dtype, device=weight.device)
            w[:, :cin, :cout] = weight
        else:
            w = weight.contiguous()
        return w.view(K, 1, cin_p, cout_p), cin_p, cout_p, cin, cout
    K, G, cin_g, cout_g = weight.shape
    assert G == groups
    if cin_g % 8 or cout_g % 8:
        raise ValueError(
            f"group conv needs channels-per-group that are multiples of 8, got {cin_g}->{cout_g} "
            f"(the reference has the same floor: detail/dispatch.py:42-50)")
    return weight.contiguous(), cin_g, cout_g, cin_g, cout_g


def _fusable_weight(weight: Tensor, compute_dtype: torch.dtype, groups: int):
    """(K, G, cin_g, cout_g) when the weight can go through wcn_weight_image_pair as it is: no
    channel padding, master dtype equal to the compute dtype or fp32 above a 16-bit one."""
    if weight.dtype != compute_dtype and not (weight.dtype == torch.float32
                                              and compute_dtype in (torch.bfloat16, torch.float16)):
        return None
    if groups == 1:
        if weight.dim() != 3:
            return None
        K, cin, cout = weight.shape
        if cin % _CH_ALIGN or cout % _CH_ALIGN:
            return None
        return K, 1, cin, cout
    if weight.dim() != 4 or weight.shape[1] != groups:
        return None
    K, G, cin_g, cout_g = weight.shape
    if cin_g % 8 or cout_g % 8:
        return None
    return K, G, cin_g, cout_g


def sparse_conv_forward(in_features: Tensor, weight: Tensor, kernel_map: IntSearchResult,
                        num_out_coords: int, groups: int = 1, bias: Optional[Tensor] = None,
                        relu: bool = False, stats: Optional[Tensor] = None) -> Tensor:
    """Y[M, Cout] = sum_k X[in_k] @ W_k  on the tensor cores (no autograd)."""
    w4, cin_g, cout_g, cin_r, cout_r = _canon_weight(weight, groups)
    K = w4.shape[0]
    x = _pad_cols(in_features, groups * cin_g)
    plan = kernel_map.fwd_plan(num_out_coords)
    img = _ops.weight_image(w4, K, groups, cin_g, cout_g, transpose_w=False)
    if bias is not None and cout_g != cout_r:
        bias = torch.nn.functional.pad(bias.float(), (0, cout_g - cout_r))
    pad_stats = None
    if stats is not None and cout_g != cout_r:  # padded channels: statistics into a padded buffer
        pad_stats = torch.zeros((2, groups * cout_g), dtype=torch.float64, device=x.device)
    y = _ops.gather_gemm(x, img, plan, groups, cin_g, cout_g,
                         bias=None if bias is None else bias.float().contiguous(), relu=relu,
                         stats=stats if pad_stats is None else pad_stats)
    if pad_stats is not None:
        stats.copy_(pad_stats[:, :cout_r])
    return y if cout_g == cout_r else y[:, :cout_r]


def sparse_conv_dgrad(grad_output: Tensor, weight: Tensor, kernel_map: IntSearchResult,
                      num_in_coords: int, groups: int = 1) -> Tensor:
    """dX[N, Cin] = sum_k dY[out_k] @ W_k^T."""
    w4, cin_g, cout_g, cin_r, cout_r = _canon_weight(weight, groups)
    K = w4.shape[0]
    gy = _pad_cols(grad_output, groups * cout_g)
    plan, kflip = kernel_map.bwd_plan(num_in_coords)
    img = _ops.weight_image(w4, K, groups, cin_g, cout_g, transpose_w=True)
    # roles swap: contraction over cout_g, rows produced = cin_g
    dx = _ops.gather_gemm(gy, img, plan, groups, cout_g, cin_g, kflip=kflip)
    return dx if cin_g == cin_r else dx[:, :cin_r]


# wgrad walks the pair lists in row-block-major order (csrc/conv_wgrad.cu: WgSegCursor) once the
# gathered operands no longer fit L2 comfortably: _WGRAD_ROUNDS windows of the rows, each seeing
# all K offsets while L2-resident. Measured on C3-S (ncu, profiles/r1d): 2 windows cut the DRAM
# reads from 345 MB to 123 MB (algorithmic: 126 MB) for +2 % kernel time (every extra segment
# restarts the index prefetch and flushes an accumulator; 4 windows: 119 MB, +7 %). Below the
# threshold the plain offset-major order is kept.
_WGRAD_LOCALITY_BYTES = 48 << 20
_WGRAD_ROUNDS = 2  # module constant: the product path reads no experiment switches from the environment


# Measured on C3-S: fetching the identity offset's rows (11 % of the pairs) as TMA tiles leaves the
# kernel time unchanged (132.1 vs 132.6 us) — the CTAs that own those pairs finish early but the
# static pair-count partition does not hand them more work, and a TMA-fed stage is still paced by
# its four 128x128x16 MMAs (~600 of the 1 140 cycles of a gathered stage). Kept behind the C-ABI,
# off by default (the two tensor-map encodes also cost host time per call).
_WGRAD_IDENTITY_TMA = False


def _wgrad_identity(kernel_map, K: int):
    """The centre offset of a submanifold map (same coordinates, odd kernel, stride 1) pairs every
    row with itself, in order: wgrad can fetch those rows as TMA tiles. The hash table's status
    word tells the kernel when duplicate coordinates break that property."""
    if not _WGRAD_IDENTITY_TMA or not getattr(kernel_map, "_symmetric", False) or K % 2 == 0:
        return {}
    table = getattr(kernel_map, "_hashtable", None)
    return {"identity_k": K // 2, "status": None if table is None else table.status_tensor}


def _wgrad_order(x: Tensor, gy: Tensor, kernel_map, K: int):
    """Row-block-major unit order once the operands exceed the L2-comfortable size.
    (A "dense rows" form — dY of high-occupancy offsets as TMA tiles, X gathered through the pair
    table — was built and measured SLOWER, 152.6 -> 170.9 us on C3-S: the kernel is bound by the
    bytes crossing L2 -> SM whichever unit requests them. Removed again; the measurement and the
    commit that holds the code are in profiles/r2_wgrad_dense_rows_negative_result.md.)"""
    bp = getattr(kernel_map, "_block_prefix", None)
    work = x.numel() * x.element_size() + gy.numel() * gy.element_size()
    if (bp is None or _WGRAD_ROUNDS <= 1 or work < _WGRAD_LOCALITY_BYTES
            or _WGRAD_ROUNDS * K > 1024 or bp.shape[0] != K or bp.shape[1] < _WGRAD_ROUNDS):
        return {}
    return {"row_block_prefix": bp, "row_parts": _WGRAD_ROUNDS, "rounds": _WGRAD_ROUNDS}


def _wgrad_identity(kernel_map, K: int):
    """The centre offset of a submanifold map (same coordinates, odd kernel, stride 1) pairs every
    row with itself, in order: wgrad can fetch those rows as TMA tiles. The hash table's status
    word tells the kernel when duplicate coordinates break that property."""
    if not _WGRAD_IDENTITY_TMA or not getattr(kernel_map, "_symmetric", False) or K % 2 == 0:
        return {}
    table = getattr(kernel_map, "_hashtable", None)
    return {"identity_k": K // 2, "status": None if table is None else table.status_tensor}


# Dense-row wgrad (csrc/conv_wgrad.cu): offsets that pair >= 60 % of the output rows are walked over
# ALL output rows in row order — dY arrives as dense TMA tiles, only X is gathered through the pair
# table — which halves the LSU-gathered bytes of those offsets. MEASURED SLOWER and therefore OFF
# (profiles/r2_wgrad_dense_rows_negative_result.md): C3-S 152.6 -> 170.9 us, C3-R 143.7 -> 159.3 us
# against the same unit order without it. A dense stage (64 rows, X gathered + dY by TMA) costs
# ~0.86 of a pair-list stage, not the ~0.54 its LSU bytes suggest: what bounds the kernel is the
# bytes that cross L2 -> SM, whichever unit requests them, and the TMA tiles move the same bytes
# (plus the rows of missing pairs). Kept behind the C-ABI, parity-tested, for the record.
_WGRAD_DENSE_ROWS = False
_WGRAD_DENSE_ROUNDS = 2   # chunks per CTA of the weighted virtual list (evens out dense / sparse)


def _wgrad_order(x: Tensor, gy: Tensor, kernel_map, K: int):
    bp = getattr(kernel_map, "_block_prefix", None)
    if bp is None or bp.shape[0] != K or K > 256:
        return {}
    work = x.numel() * x.element_size() + gy.numel() * gy.element_size()
    big = (_WGRAD_ROUNDS > 1 and work >= _WGRAD_LOCALITY_BYTES and _WGRAD_ROUNDS * K <= 1024
           and bp.shape[1] >= _WGRAD_ROUNDS)
    order = {"row_block_prefix": bp, "row_parts": _WGRAD_ROUNDS, "rounds": _WGRAD_ROUNDS} if big else {}
    pt = getattr(kernel_map, "_pair_table", None)
    if (_WGRAD_DENSE_ROWS and pt is not None and pt.shape == (K, gy.shape[0])
            and x.dtype in (torch.bfloat16, torch.float16) and gy.shape[1] * 2 <= 256
            and x.shape[1] * 2 <= 256 and bp.shape[1] * 256 >= gy.shape[0]):
        if not order:
            order = {"row_block_prefix": bp, "row_parts": 1, "rounds": _WGRAD_DENSE_ROUNDS}
        order["pair_table"] = pt
    return order


def _wgrad_call(x: Tensor, gy: Tensor, kernel_map, K: int, G: int, cin_g: int, cout_g: int,
                out: Optional[Tensor] = None) -> Tensor:
    im, om, od = kernel_map._in_buf, kernel_map._out_buf, kernel_map.offsets_dev
    order = dict(_wgrad_order(x, gy, kernel_map, K), **_wgrad_identity(kernel_map, K))
    if out is not None:  # caller's buffer (e.g. a peer-mapped all-reduce buffer): zero, then reduce into it
        assert out.dtype == torch.float32 and out.is_contiguous() \
            and out.numel() == K * G * cin_g * cout_g
        order["dw"] = out.zero_().view(K, G, cin_g, cout_g)
    if x.dtype != torch.float32:
        return _ops.wgrad(x, gy, im, om, od, K, G, cin_g, cout_g, **order)
    # fp32 operands: the contraction runs over gathered rows (MN-major operands), which the
    # tensor cores take as 16-bit types only, so each operand is split into bf16 hi + lo parts and
    # three bf16 products (hi*hi + hi*lo + lo*hi, fp32 accumulate) are summed into one dW:
    # ~16 mantissa bits, tighter than a single TF32 pass. (The reference downcasts fp32 inputs of
    # its mask_gemm path to fp16 outright, detail/mask_gemm.py:65-103.)
    xh = x.bfloat16()
    gh = gy.bfloat16()
    xl = (x - xh.float()).bfloat16()
    gl = (gy - gh.float()).bfloat16()
    dw = _ops.wgrad(xh, gh, im, om, od, K, G, cin_g, cout_g, **order)
    order["dw"] = dw
    _ops.wgrad(xh, gl, im, om, od, K, G, cin_g, cout_g, **order)
    _ops.wgrad(xl, gh, im, om, od, K, G, cin_g, cout_g, **order)
    return dw


def sparse_conv_wgrad(in_features: Tensor, grad_output: Tensor, weight_shape, kernel_map,
                      groups: int = 1, out: Optional[Tensor] = None) -> Tensor:
    """fp32 dW with the shape of the weight. ``out`` (fp32, contiguous, the weight's element
    count, channel counts that need no padding) receives the result in place of a fresh tensor —
    e.g. a ``PeerAllReduce`` buffer, so the reduction needs no packing copy."""
    if groups == 1:
        K, cin, cout = weight_shape
        cin_p, cout_p = _round_up(cin, _CH_ALIGN), _round_up(cout, _CH_ALIGN)
        if out is not None and (cin_p != cin or cout_p != cout):
            raise ValueError("out= needs channel counts that are multiples of 16")
        x = _pad_cols(in_features, cin_p)
        gy = _pad_cols(grad_output, cout_p)
        dw = _wgrad_call(x, gy, kernel_map, K, 1, cin_p, cout_p, out=out).view(K, cin_p, cout_p)
        return dw if (cin_p == cin and cout_p == cout) else dw[:, :cin, :cout]
    K, G, cin_g, cout_g = weight_shape
    x = _pad_cols(in_features, G * cin_g)
    gy = _pad_cols(grad_output, G * cout_g)
    return _wgrad_call(x, gy, kernel_map, K, G, cin_g, cout_g, out=out)


class UnifiedSpatiallySparseConvFunction(Function):
    """Same positional signature as the reference (unified.py:145-166) so call sites match."""

    @staticmethod
    def forward(ctx, in_features, weight, kernel_map, num_out_coords, fwd_algo=None,
                dgrad_algo=None, wgrad_algo=None, compute_dtype=None, fwd_block_size=None,
                bwd_block_size=None, in_tensor_stride=None, conv_cache_metadata=None, groups=1,
                use_fp16_accum=False, bias=None, stats=None):
        """``bias`` (fp32 [Cout]) is added in the GEMM epilogue; ``stats`` (zero-filled fp64
        [2, Cout]) receives the per-channel sum / sum of squares of the stored output."""
        if not in_features.is_cuda:
            raise RuntimeError("warpconvnet_b200 sparse conv needs CUDA tensors (no CPU fallback)")
        out_dtype = in_features.dtype
        x = in_features
        if compute_dtype is not None and x.dtype != compute_dtype:
            x = x.to(compute_dtype)
        # One launch prepares the forward AND the dgrad weight image straight from the master
        # weights (fp32 -> compute dtype inside the kernel): no weight.to(bf16) copy, no second
        # image launch in backward. Needs channel counts that take no padding.
        fused = _fusable_weight(weight, x.dtype, groups)
        if fused:
            K, G, cin_g, cout_g = fused
            w4 = weight.detach().contiguous().view(K, G, cin_g, cout_g)
            img, img_t = _ops.weight_image_pair(w4, K, G, cin_g, cout_g, x.dtype,
                                                want_transposed=bool(ctx.needs_input_grad[0]))
            xp = _pad_cols(x, G * cin_g)
            y = _ops.gather_gemm(xp, img, kernel_map.fwd_plan(num_out_coords), G, cin_g, cout_g,
                                 bias=None if bias is None else bias.detach().float().contiguous(),
                                 stats=stats)
            ctx.save_for_backward(x, img_t)
            ctx.fused_dims = (G, cin_g, cout_g)
        else:
            w = weight if weight.dtype == x.dtype else weight.to(x.dtype)
            y = sparse_conv_forward(x, w, kernel_map, num_out_coords, groups,
                                    bias=None if bias is None else bias.detach(), stats=stats)
            ctx.save_for_backward(x, w)
            ctx.fused_dims = None
        ctx.kernel_map = kernel_map
        ctx.groups = groups
        ctx.in_dtype = in_features.dtype
        ctx.w_dtype = weight.dtype
        ctx.num_in = in_features.shape[0]
        ctx.weight_shape = tuple(weight.shape)
        ctx.bias_dtype = None if bias is None else bias.dtype
        return y if compute_dtype is None else y.to(out_dtype)

    @staticmethod
    def backward(ctx, grad_output):
        x, w = ctx.saved_tensors   # fused path: w is the dgrad weight image (or None)
        kernel_map = ctx.kernel_map
        if not torch.cuda.is_current_stream_capturing():
            check_pending_kernel_maps()  # deferred coordinate-range / table-full errors
        gy = grad_output
        if gy.dtype != x.dtype:
            gy = gy.to(x.dtype)
        if not gy.is_contiguous():
            gy = gy.contiguous()
        grad_in = grad_w = None
        if ctx.needs_input_grad[0] and ctx.fused_dims is not None:
            G, cin_g, cout_g = ctx.fused_dims
            plan, kflip = kernel_map.bwd_plan(ctx.num_in)
            grad_in = _ops.gather_gemm(_pad_cols(gy, G * cout_g), w, plan, G, cout_g, cin_g,
                                       kflip=kflip).to(ctx.in_dtype)
        elif ctx.needs_input_grad[0]:
            grad_in = sparse_conv_dgrad(gy, w, kernel_map, ctx.num_in, ctx.groups).to(ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            grad_w = sparse_conv_wgrad(x, gy, ctx.weight_shape, kernel_map, ctx.groups)
            grad_w = grad_w.to(ctx.w_dtype)
            if grad_w.shape != ctx.weight_shape:
                grad_w = grad_w.reshape(ctx.weight_shape)
        grad_b = None
        if ctx.bias_dtype is not None and ctx.needs_input_grad[14]:
            grad_b = grad_output.sum(dim=0, dtype=torch.float32).to(ctx.bias_dtype)
        return (grad_in, grad_w) + (None,) * 12 + (grad_b, None)
