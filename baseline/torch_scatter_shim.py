# SPDX-License-Identifier: Apache-2.0
"""Minimal stand-in for ``torch_scatter`` (not installed in this image) so that the UNMODIFIED
reference package under baseline/_ref imports. Only ``segment_csr`` is used by the reference
(warpconvnet/geometry/coords/ops/reductions.py, nn/functional/normalizations.py); it is written
here with plain torch ops and is NOT on the timed sparse-conv path of the head-to-head bench."""
import torch


def segment_csr(src: torch.Tensor, indptr: torch.Tensor, out=None, reduce: str = "sum"):
    indptr = indptr.to(src.device).long()
    n = indptr.numel() - 1
    counts = indptr[1:] - indptr[:-1]
    seg = torch.repeat_interleave(torch.arange(n, device=src.device), counts)
    shape = (n,) + tuple(src.shape[1:])
    red = {"sum": "sum", "add": "sum", "mean": "mean", "max": "amax", "min": "amin"}[reduce]
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    idx = seg.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    res = res.scatter_reduce(0, idx, src, reduce=red, include_self=False)
    if out is not None:
        out.copy_(res)
        return out
    return res


def segment_coo(src, index, out=None, dim_size=None, reduce="sum"):
    n = int(index.max()) + 1 if dim_size is None else dim_size
    shape = (n,) + tuple(src.shape[1:])
    red = {"sum": "sum", "add": "sum", "mean": "mean", "max": "amax", "min": "amin"}[reduce]
    idx = index.long().view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_reduce(
        0, idx, src, reduce=red, include_self=False)


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0
    return segment_coo(src, index, out, dim_size, reduce)
