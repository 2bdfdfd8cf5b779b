# SPDX-License-Identifier: Apache-2.0
"""``spatially_sparse_conv`` — the functional entry point behind ``SparseConv3d``
(drop-in for warpconvnet/nn/functional/sparse_conv/helper.py:147-567: same arguments, same
output-coordinate rules, same kernel-map caching on the ``Voxels`` object, same transposed-map
reuse)."""
from __future__ import annotations

from enum import Enum
from typing import List, Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from warpconvnet_b200.geometry.base.geometry import Geometry
from warpconvnet_b200.geometry.coords.integer import IntCoords
from warpconvnet_b200.geometry.coords.ops.stride import stride_coords
from warpconvnet_b200.geometry.coords.search.cache import IntSearchCache, IntSearchCacheKey
from warpconvnet_b200.geometry.coords.search.search_results import IntSearchResult
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map
from warpconvnet_b200.geometry.types.voxels import Voxels
from warpconvnet_b200.utils.ntuple import ntuple

from .detail.unified import (SPARSE_CONV_AB_ALGO_MODE, SPARSE_CONV_ATB_ALGO_MODE,
                             UnifiedSpatiallySparseConvFunction)


class STRIDED_CONV_MODE(Enum):
    REDUCE_AND_STRIDE = "reduce_and_stride"
    STRIDE_ONLY = "stride_only"


def _intcoords_from_batch_indexed(reference: IntCoords, bcoords: Tensor, offsets: Tensor):
    return reference.__class__(bcoords[:, 1:].contiguous(),
                               offsets.to(device="cpu", dtype=reference.offsets.dtype),
                               voxel_size=reference.voxel_size,
                               tensor_stride=reference.tensor_stride)


def _apply_generative_policy(input_sparse_tensor: Voxels, kernel_size, kernel_dilation, stride,
                             stride_mode, transposed):
    """Output coordinates of a generative conv (helper.py:58-144)."""
    input_coords = input_sparse_tensor.batched_coordinates
    bin_coords = input_sparse_tensor.batch_indexed_coordinates
    if all(s == 1 for s in stride):
        ex = input_coords.expand(kernel_size, kernel_dilation)
        return ex.batch_indexed_coordinates, ex.offsets, bin_coords
    if transposed:
        scaled = bin_coords.clone()  # per-column scalar multiply: no host->device tensor, no sync
        for d, s_d in enumerate(stride):
            scaled[:, d + 1] *= int(s_d)
        ex = _intcoords_from_batch_indexed(input_coords, scaled, input_sparse_tensor.offsets
                                           ).expand(kernel_size, kernel_dilation)
        return ex.batch_indexed_coordinates, ex.offsets, scaled
    strided, strided_offsets = stride_coords(bin_coords, stride,
                                             n_batches=len(input_sparse_tensor.offsets) - 1)
    strided_coords = _intcoords_from_batch_indexed(input_coords, strided, strided_offsets)
    ex = strided_coords.expand(kernel_size, kernel_dilation)
    km_in = bin_coords if stride_mode == STRIDED_CONV_MODE.STRIDE_ONLY \
        else strided_coords.batch_indexed_coordinates
    return ex.batch_indexed_coordinates, ex.offsets, km_in


@torch.compiler.disable
def spatially_sparse_conv(input_sparse_tensor, weight, kernel_size, stride=1, kernel_dilation=1,
                          bias=None, groups=1, use_fp16_accum=None, kernel_matmul_batch_size=2,
                          generative=False, output_spatially_sparse_tensor=None, transposed=False,
                          fwd_algo=SPARSE_CONV_AB_ALGO_MODE.TCGEN05,
                          dgrad_algo=SPARSE_CONV_AB_ALGO_MODE.TCGEN05,
                          wgrad_algo=SPARSE_CONV_ATB_ALGO_MODE.TCGEN05,
                          stride_mode=STRIDED_CONV_MODE.STRIDE_ONLY, stride_reduce="max", order=None,
                          compute_dtype=None, implicit_matmul_fwd_block_size=16,
                          implicit_matmul_bwd_block_size=16, bn_stats: bool = False) -> Geometry:
    """Functional sparse convolution (same keywords as the reference's, helper.py:147-358)."""
    if not isinstance(input_sparse_tensor, Voxels):
        raise TypeError("Native spatially_sparse_conv expects input_sparse_tensor of type Voxels, "
                        f"got {type(input_sparse_tensor)}")
    if output_spatially_sparse_tensor is not None and not isinstance(
            output_spatially_sparse_tensor, Voxels):
        raise TypeError("Native spatially_sparse_conv expects output_spatially_sparse_tensor of "
                        f"type Voxels or None, got {type(output_spatially_sparse_tensor)}")

    nd = input_sparse_tensor.num_spatial_dims
    _kernel_size = ntuple(kernel_size, ndim=nd)
    _kernel_dilation = ntuple(kernel_dilation, ndim=nd)
    _stride = ntuple(stride, ndim=nd)

    # 1x1x1 stride-1 shortcut: plain dense matmul, no kernel map (helper.py:206-213)
    if np.prod(_kernel_size) == 1 and np.prod(_stride) == 1:
        feats = input_sparse_tensor.feature_tensor
        w0 = weight[0]
        if w0.dim() == 3:  # grouped [G, cin_g, cout_g] -> block diagonal
            w0 = torch.block_diag(*w0.unbind(0))
        if bias is not None:  # bias in the GEMM epilogue (one pass less over the output)
            out = torch.addmm(bias.to(feats.dtype), feats, w0.to(feats.dtype))
        else:
            out = feats @ w0.to(feats.dtype)
        return input_sparse_tensor.replace(batched_features=out)

    in_tensor_stride = input_sparse_tensor.tensor_stride
    if in_tensor_stride is None:
        in_tensor_stride = ntuple(1, ndim=nd)
    if transposed and not generative:
        assert output_spatially_sparse_tensor is not None, \
            "Output spatially sparse tensor is required for transposed convolution without generative"

    if not transposed:
        out_tensor_stride = tuple(o * s for o, s in zip(_stride, in_tensor_stride))
    elif generative:
        out_tensor_stride = tuple(i // s for i, s in zip(in_tensor_stride, _stride))
    else:
        if output_spatially_sparse_tensor.tensor_stride is not None:
            out_tensor_stride = output_spatially_sparse_tensor.tensor_stride
        else:
            out_tensor_stride = ntuple(1, ndim=nd)
        assert any(o < i for o, i in zip(out_tensor_stride, in_tensor_stride)), \
            "Output stride is larger than input stride"

    effective_compute_dtype = compute_dtype
    if effective_compute_dtype is None:
        effective_compute_dtype = (torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled()
                                   else input_sparse_tensor.batched_features.dtype)

    if stride_mode == STRIDED_CONV_MODE.REDUCE_AND_STRIDE and any(s != 1 for s in _stride):
        # pool the input over stride-sized windows, then convolve the pooled voxels at stride 1
        # (helper.py:275-288); the kernel map below indexes the pooled rows
        from warpconvnet_b200.nn.functional.sparse_pool import sparse_reduce
        input_sparse_tensor = sparse_reduce(input_sparse_tensor, kernel_size=_stride,
                                            stride=_stride, reduction=stride_reduce)
        _stride = ntuple(1, ndim=nd)

    feats = input_sparse_tensor.batched_features.batched_tensor
    bout, out_offsets, kernel_map = generate_output_coords_and_kernel_map(
        input_sparse_tensor=input_sparse_tensor, kernel_size=_kernel_size,
        kernel_dilation=_kernel_dilation, stride=_stride, generative=generative,
        transposed=transposed, output_spatially_sparse_tensor=output_spatially_sparse_tensor,
        stride_mode=stride_mode, order=order)
    num_out = bout.shape[0]

    # Features are cast here (like the reference, helper.py:310-320); the weight goes into the
    # autograd function in its master dtype and is cast inside, so the fp32 wgrad result reaches
    # weight.grad without a round trip through bf16.
    x, w = feats, weight
    if x.dtype != effective_compute_dtype:
        x = x.to(effective_compute_dtype)
    if groups > 1 and (weight.shape[2] % 8 or weight.shape[3] % 8):
        # fewer than 8 channels per group (the reference's mask_gemm path rejects these outright,
        # detail/dispatch.py:42-50): run the DENSE tensor-core kernels on the block-diagonal
        # [K, Cin, Cout] weight. The off-diagonal zeros cost G x the useful FLOPs, on matrices this
        # narrow the kernels stay gather-bound; autograd returns the diagonal blocks of dW.
        eye = torch.eye(groups, dtype=weight.dtype, device=weight.device)
        kk, gg, cig, cog = weight.shape
        w = torch.einsum("kgio,gh->kgiho", weight, eye).reshape(kk, gg * cig, gg * cog)
        groups = 1

    # bias (and, on request, the statistics of the BatchNorm that follows) go through the GEMM
    # epilogue: no separate add pass over Y, no separate statistics pass
    stats = None
    if bn_stats and num_out > 0:
        stats = torch.zeros((2, w.shape[-1] * (groups if w.dim() == 4 else 1)),
                            dtype=torch.float64, device=x.device)
    out = UnifiedSpatiallySparseConvFunction.apply(
        x, w, kernel_map, num_out, fwd_algo, dgrad_algo, wgrad_algo, effective_compute_dtype,
        implicit_matmul_fwd_block_size, implicit_matmul_bwd_block_size, in_tensor_stride, None,
        groups, bool(use_fp16_accum), bias, stats)
    if stats is not None:
        out._wcn_bn_sums = (stats, out.shape[0], out._version)

    out_offsets_cpu = out_offsets if out_offsets.device.type == "cpu" else out_offsets.cpu()
    if bout is input_sparse_tensor.batch_indexed_coordinates:
        # submanifold: same coordinate object -> keep it (and its cached [N,4] view)
        new_coords = input_sparse_tensor.batched_coordinates
        if tuple(out_tensor_stride) != tuple(in_tensor_stride):
            new_coords = IntCoords(new_coords.batched_tensor, offsets=out_offsets_cpu)
    else:
        new_coords = IntCoords(bout[:, 1:].contiguous(), offsets=out_offsets_cpu)
        new_coords._bcoords = bout
    return input_sparse_tensor.replace(batched_coordinates=new_coords, batched_features=out,
                                       tensor_stride=out_tensor_stride)


@torch.compiler.disable
def generate_output_coords_and_kernel_map(input_sparse_tensor, kernel_size, kernel_dilation, stride,
                                          generative=False, transposed=False,
                                          output_spatially_sparse_tensor=None,
                                          stride_mode=STRIDED_CONV_MODE.STRIDE_ONLY, order=None,
                                          kernel_search_batch_size=None, out_code_backend=None):
    """(batch-indexed output coordinates, output offsets, kernel map) for one conv; kernel maps are
    cached on the input ``Voxels`` and transposed convs reuse the encoder's map with in / out
    swapped (same rules as the reference, helper.py:361-567)."""
    bin_coords = input_sparse_tensor.batch_indexed_coordinates
    same_coords = False
    if output_spatially_sparse_tensor is not None:
        assert not generative, \
            "Output spatially sparse tensor is not supported with generative convolution"
        bout = output_spatially_sparse_tensor.batch_indexed_coordinates
        out_offsets = output_spatially_sparse_tensor.offsets
    elif generative:
        bout, out_offsets, bin_coords = _apply_generative_policy(
            input_sparse_tensor, kernel_size, kernel_dilation, stride, stride_mode, transposed)
    elif any(s != 1 for s in stride):
        # Strided coordinates are memoised on the input's IntCoords OBJECT (not in the
        # _extra_attributes dict that replace() shares with every descendant tensor): the object
        # is replaced whenever the coordinates change, so a later level with the same per-scene
        # counts can never be handed this level's result, and .to(device) / sort / unique (which
        # build new IntCoords) drop it. The reference recomputes stride_coords on every call.
        cobj = input_sparse_tensor.batched_coordinates
        cache = cobj.__dict__.setdefault("_stride_cache", {})
        key = (tuple(stride), bin_coords.data_ptr(), bin_coords.shape[0])
        if key not in cache:
            cache[key] = stride_coords(bin_coords, stride,
                                       n_batches=len(input_sparse_tensor.offsets) - 1)
        bout, out_offsets = cache[key]
    else:
        bout, out_offsets = bin_coords, input_sparse_tensor.offsets
        same_coords = True

    # Output-row ordering (helper.py:435-442): any ordering but RANDOM re-sorts the output
    # coordinates of every batch item along the requested space-filling curve before the map is
    # built (the ordered rows are no longer the input's rows, so no submanifold shortcuts).
    if order is not None:
        from warpconvnet_b200.geometry.coords.ops.serialization import (POINT_ORDERING,
                                                                        STR2POINT_ORDERING, encode)
        if isinstance(order, str):
            order = STR2POINT_ORDERING[order.lower()]
        if order is not POINT_ORDERING.RANDOM:
            perm = encode(bout[:, 1:], batch_offsets=out_offsets, order=order,
                          return_perm=True).perm
            bout = bout[perm].contiguous()
            same_coords = False

    key = IntSearchCacheKey(kernel_size=kernel_size, kernel_dilation=kernel_dilation,
                            transposed=transposed, generative=generative,
                            stride_mode=str(stride_mode), skip_symmetric_kernel_map=False,
                            in_offsets=input_sparse_tensor.offsets, out_offsets=out_offsets)
    if input_sparse_tensor.cache is not None and isinstance(input_sparse_tensor.cache,
                                                            IntSearchCache):
        hit = input_sparse_tensor.cache.get(key)
        if hit is not None:
            return bout, out_offsets, hit

    if transposed and not generative:
        fwd_key = IntSearchCacheKey(kernel_size=kernel_size, kernel_dilation=kernel_dilation,
                                    transposed=False, generative=generative,
                                    stride_mode=str(stride_mode), skip_symmetric_kernel_map=False,
                                    in_offsets=out_offsets,
                                    out_offsets=input_sparse_tensor.offsets)
        fwd_map = None
        for src in (input_sparse_tensor, output_spatially_sparse_tensor):
            if src is not None and isinstance(src.cache, IntSearchCache):
                fwd_map = src.cache.get(fwd_key)
                if fwd_map is not None:
                    break
        if fwd_map is None:
            fwd_map = generate_kernel_map(bout, bin_coords, stride, kernel_size, kernel_dilation)
        kernel_map = fwd_map.transposed_view()
    elif transposed and generative:
        kernel_map = generate_kernel_map(
            bout, bin_coords, ntuple(1, ndim=input_sparse_tensor.num_spatial_dims), kernel_size,
            kernel_dilation).transposed_view()
    elif stride_mode == STRIDED_CONV_MODE.STRIDE_ONLY or all(s == 1 for s in stride):
        # REDUCE_AND_STRIDE arrives here with the pooled voxels and stride 1 (helper.py:539-558:
        # out-to-out map, or in-to-expanded map when generative)
        kernel_map = generate_kernel_map(bin_coords, bout, stride, kernel_size, kernel_dilation,
                                         same_coords=same_coords)
    else:
        raise ValueError(f"Unsupported case. stride_mode: {stride_mode}, generative: {generative}, "
                         f"transposed: {transposed}")

    if not isinstance(input_sparse_tensor.cache, IntSearchCache):
        input_sparse_tensor._extra_attributes["_cache"] = IntSearchCache()
    input_sparse_tensor.cache.put(key, kernel_map)
    return bout, out_offsets, kernel_map
