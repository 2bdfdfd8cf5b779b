# SPDX-License-Identifier: Apache-2.0
"""GPU parity of the fused BatchNorm (+ residual) (+ ReLU) row kernels (csrc/rownorm.cu, through the
C-ABI) against torch's own batch_norm in fp64 on the CPU — the op the reference's ConvBlock /
BasicBlock apply between sparse convs (models/mink_unet.py:31-53,104-140)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {torch.bfloat16: 2e-2, torch.float16: 3e-3, torch.float32: 2e-5}


def _ref(x, w, b, res, relu, training, rm, rv, eps=1e-5, momentum=0.1):
    x = x.detach().double().cpu().requires_grad_(True)
    w = w.detach().double().cpu().requires_grad_(True)
    b = b.detach().double().cpu().requires_grad_(True)
    res = res.detach().double().cpu().requires_grad_(True) if res is not None else None
    rm, rv = rm.double().cpu().clone(), rv.double().cpu().clone()
    y = torch.nn.functional.batch_norm(x, rm, rv, w, b, training, momentum, eps)
    if res is not None:
        y = y + res
    if relu:
        y = torch.relu(y)
    return x, w, b, res, y, rm, rv


def _rel(a, ref):
    return float((a.detach().double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("n,c", [(4099, 32), (1000, 96), (777, 20), (2500, 256), (1, 64), (513, 3)])
@pytest.mark.parametrize("relu,with_res", [(False, False), (True, False), (True, True)])
def test_batch_norm_act_training(dtype, n, c, relu, with_res):
    from warpconvnet_b200.nn.functional.normalizations import batch_norm_act
    g = torch.Generator().manual_seed(n * 131 + c)
    x = (torch.randn(n, c, generator=g) * 1.7 + 0.4).to(dtype).cuda().requires_grad_(True)
    w = (torch.rand(c, generator=g) + 0.5).cuda().requires_grad_(True)
    b = torch.randn(c, generator=g).cuda().requires_grad_(True)
    res = torch.randn(n, c, generator=g).to(dtype).cuda().requires_grad_(True) if with_res else None
    rm, rv = torch.zeros(c).cuda(), torch.ones(c).cuda()
    gy = torch.randn(n, c, generator=g).to(dtype).cuda()
    if n == 1:
        # nn.BatchNorm1d refuses a single row in training mode ("Expected more than 1 value per
        # channel when training": the unbiased running variance divides by n - 1); so does ours
        with pytest.raises(ValueError):
            batch_norm_act(x, w, b, rm, rv, training=True, relu=False, residual=None)
        return
    rx, rw, rb, rres, ry, rrm, rrv = _ref(x, w, b, res, relu, True, rm, rv)
    y = batch_norm_act(x, w, b, rm, rv, training=True, relu=relu, residual=res)
    assert y.dtype == dtype and y.shape == (n, c)
    tol = TOL[dtype]
    assert _rel(y, ry.detach()) < tol
    assert _rel(rm, rrm) < 1e-4 and _rel(rv, rrv) < 1e-4       # running statistics (fp32)
    # the mask of the oracle is taken on ITS output; use a gradient that vanishes where the two
    # could disagree (|y| tiny) so the comparison is about arithmetic, not about rounding of zeros
    gyd = gy.double().cpu() * (ry.detach().abs() > 0.05 if relu else 1.0)
    ry.backward(gyd)
    y.backward(gyd.to(dtype).cuda())
    assert _rel(x.grad, rx.grad) < 2 * tol
    assert _rel(w.grad, rw.grad) < 2 * tol
    assert _rel(b.grad, rb.grad) < 2 * tol
    if with_res:
        assert _rel(res.grad, rres.grad) < 2 * tol


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_batch_norm_eval_and_module(dtype):
    from warpconvnet_b200.nn.modules.normalizations import BatchNorm
    n, c = 3000, 64
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, c, generator=g).to(dtype).cuda()
    ours = BatchNorm(c, relu=True).cuda()
    ref = torch.nn.BatchNorm1d(c).double()
    with torch.no_grad():
        ours.norm.weight.copy_(torch.rand(c, generator=g) + 0.5)
        ours.norm.bias.copy_(torch.randn(c, generator=g))
    ref.load_state_dict({k: v.double().cpu() if v.is_floating_point() else v.cpu()
                         for k, v in ours.norm.state_dict().items()})
    for _ in range(3):                                  # training steps move the running stats
        y = ours(x)
        ry = torch.relu(ref(x.double().cpu()))
        assert _rel(y, ry.detach()) < TOL[dtype]
    assert int(ours.norm.num_batches_tracked) == 3
    assert _rel(ours.norm.running_mean, ref.running_mean) < 1e-4
    assert _rel(ours.norm.running_var, ref.running_var) < 1e-4
    ours.eval(); ref.eval()
    xg = x.clone().requires_grad_(True)
    xr = x.double().cpu().requires_grad_(True)
    y = ours(xg)
    ry = torch.relu(ref(xr))
    assert _rel(y, ry.detach()) < TOL[dtype]
    gy = torch.randn(n, c, generator=g)
    gyd = gy.double() * (ry.detach().abs() > 0.05)
    ry.backward(gyd)
    y.backward(gyd.to(dtype).cuda())
    assert _rel(xg.grad, xr.grad) < 2 * TOL[dtype]


def test_batch_norm_rejects_cpu():
    from warpconvnet_b200.nn.functional.normalizations import batch_norm_act
    with pytest.raises(RuntimeError):
        batch_norm_act(torch.randn(8, 8))
