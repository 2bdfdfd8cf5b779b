# SPDX-License-Identifier: Apache-2.0
"""Fused BatchNorm (+ residual) (+ ReLU) on the feature matrix between two sparse convolutions.

Replaces the torch chain the reference runs per ConvBlock / BasicBlock
(``nn.BatchNorm1d -> ReLU`` and ``bn(x) + identity -> ReLU``; warpconvnet/models/mink_unet.py:31-53,
104-140; warpconvnet/nn/modules/normalizations.py:53-67) with the row-streaming kernels of
``csrc/rownorm.cu``: forward = statistics pass + one normalise/activate pass, backward = one
reduction pass + one apply pass. Semantics are nn.BatchNorm1d's (biased variance for the
normalisation, unbiased for ``running_var``, fp32 statistics whatever the feature dtype).
"""
from typing import Optional

import torch
from torch import Tensor

from warpconvnet_b200 import _ops


def _fp32_stat(t: Optional[Tensor], c: int, device) -> Optional[Tensor]:
    """The tensor itself when the native kernel may update it in place (fp32, contiguous, [c], on
    the feature device), else an fp32 copy the caller writes back."""
    if t is None:
        return None
    if t.numel() != c:
        raise ValueError(f"running statistic has {t.numel()} entries for {c} channels")
    if t.device != device:
        raise ValueError(f"running statistic lives on {t.device}, features on {device}")
    if t.dtype == torch.float32 and t.is_contiguous():
        return t
    return t.detach().float().contiguous()


class _BatchNormAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Optional[Tensor], bias: Optional[Tensor],
                residual: Optional[Tensor], running_mean: Optional[Tensor],
                running_var: Optional[Tensor], training: bool, momentum: float, eps: float,
                relu: bool, sums: Optional[Tensor] = None):
        x = x if x.stride(1) == 1 else x.contiguous()
        n, c = x.shape
        gamma = weight.detach().float().contiguous() if weight is not None else None
        beta = bias.detach().float().contiguous() if bias is not None else None
        if residual is not None:
            residual = residual.to(x.dtype)
            residual = residual if residual.stride(1) == 1 else residual.contiguous()
        use_batch = training or running_mean is None
        if use_batch:
            if n <= 1:
                # nn.BatchNorm1d: "Expected more than 1 value per channel when training" (the
                # unbiased running variance divides by n - 1)
                raise ValueError(f"BatchNorm in training mode needs more than one row, got {n}")
            rm = running_mean if training else None
            rv = running_var if training else None
            # the native kernel reads and writes the running statistics as fp32: after
            # model.half() / .bfloat16() the nn.BatchNorm1d buffers are 2-byte types, so update
            # fp32 temporaries and copy them back (never hand a narrower buffer to the kernel)
            rm32, rv32 = _fp32_stat(rm, c, x.device), _fp32_stat(rv, c, x.device)
            y, scale, shift, mean_rstd = _ops.bn_forward(
                x, gamma, beta, eps, momentum, rm32, rv32, residual, relu, sums=sums)
            if rm32 is not rm and rm is not None:
                rm.copy_(rm32)
            if rv32 is not rv and rv is not None:
                rv.copy_(rv32)
        else:
            rstd = torch.rsqrt(running_var.float() + eps)
            scale = rstd if gamma is None else gamma * rstd
            shift = -running_mean.float() * scale
            if beta is not None:
                shift = shift + beta
            scale, shift = scale.contiguous(), shift.contiguous()
            # (mean, rstd) of the frozen statistics: backward needs them for d gamma
            mean_rstd = torch.stack([running_mean.float(), rstd]).contiguous()
            y = _ops.scale_shift_act(x, scale, shift, residual, relu)
        ctx.use_batch = use_batch
        ctx.relu = relu
        ctx.has_res = residual is not None
        ctx.has_w = weight is not None and ctx.needs_input_grad[1]
        ctx.has_b = bias is not None and ctx.needs_input_grad[2]
        ctx.w_dtype = weight.dtype if weight is not None else None
        ctx.b_dtype = bias.dtype if bias is not None else None
        # ReLU mask in backward: recomputed from x (x * scale + shift > 0) when there is no residual
        # and the output dtype keeps fp32's exponent range (bf16 / fp32), so the saved output is
        # not read again; otherwise taken from the saved output
        ctx.mask_from_x = bool(relu and use_batch and residual is None
                               and x.dtype != torch.float16)
        ctx.save_for_backward(x, y if (relu and not ctx.mask_from_x) else None,
                              gamma if gamma is not None else None, mean_rstd,
                              None if use_batch else scale,
                              scale if ctx.mask_from_x else None,
                              shift if ctx.mask_from_x else None)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, y, gamma, mean_rstd, eval_scale, msc, msh = ctx.saved_tensors
        dy = dy.to(x.dtype)
        dy = dy if dy.stride(1) == 1 else dy.contiguous()
        c = x.shape[1]
        need_res = ctx.has_res and ctx.needs_input_grad[3]
        if ctx.use_batch:
            g = gamma if gamma is not None else torch.ones(c, dtype=torch.float32, device=x.device)
            dx, dres, sums = _ops.bn_backward(dy, x, y, g, mean_rstd, msc, msh, need_res)
            dgamma = sums[1].to(ctx.w_dtype) if ctx.has_w else None
            dbeta = sums[0].to(ctx.b_dtype) if ctx.has_b else None
        else:
            # running statistics: y = x * scale + shift is affine in x
            dx, dres = _ops.bn_bwd_apply(dy, None, y, eval_scale, None, None, False, need_res)
            dgamma = dbeta = None
            if ctx.has_w or ctx.has_b:
                # frozen statistics, trainable affine (fine-tuning with BN in .eval()):
                # d beta = sum dz, d gamma = sum dz * (x - mean) * rstd, dz = dy * relu mask
                sums = _ops.bn_bwd_reduce(dy, x, y, mean_rstd)
                dgamma = sums[1].to(ctx.w_dtype) if ctx.has_w else None
                dbeta = sums[0].to(ctx.b_dtype) if ctx.has_b else None
        return dx, dgamma, dbeta, dres, None, None, None, None, None, None, None


def batch_norm_act(x: Tensor, weight: Optional[Tensor] = None, bias: Optional[Tensor] = None,
                   running_mean: Optional[Tensor] = None, running_var: Optional[Tensor] = None,
                   training: bool = True, momentum: float = 0.1, eps: float = 1e-5,
                   relu: bool = False, residual: Optional[Tensor] = None,
                   sums: Optional[Tensor] = None) -> Tensor:
    """``act(batch_norm(x) (+ residual))`` on a CUDA feature matrix ``[n, c]`` (bf16 / fp16 /
    fp32); same arguments as ``torch.nn.functional.batch_norm`` plus the fused tail. ``sums``:
    fp64 [2, c] sum / sum of squares of x from the producing conv's epilogue (training mode)."""
    if not x.is_cuda:
        raise RuntimeError("warpconvnet_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    if x.dim() != 2:
        raise ValueError(f"expected a feature matrix [n, c], got {tuple(x.shape)}")
    if sums is not None and (not training or sums.shape != (2, x.shape[1])
                             or sums.dtype != torch.float64):
        sums = None
    return _BatchNormAct.apply(x, weight, bias, residual, running_mean, running_var, training,
                               float(momentum), float(eps), bool(relu), sums)
