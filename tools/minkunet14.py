# SPDX-License-Identifier: Apache-2.0
"""MinkUNet-14-shaped network built on warpconvnet_b200.SparseConv3d — BENCHMARK / TEST HARNESS for
BASELINE config 4, not part of the package (model zoos are out of scope).

Same topology as the reference's ``MinkUNetBase(in, out, planes=(32,64,128,256,128,128,96,96),
layers=(1,)*8)`` (warpconvnet/models/mink_unet.py:259-404): 1x1 stem, four [2^3 stride-2 conv +
BasicBlock] stages, four [2^3 transposed conv + skip concat + BasicBlock] stages, 1x1 head;
BatchNorm + ReLU (+ residual add) between convs run as the fused row kernels of csrc/rownorm.cu
(the reference applies nn.BatchNorm1d / nn.ReLU / add as separate torch passes)."""
import torch
import torch.nn as nn

from warpconvnet_b200.geometry.types.voxels import Voxels
from warpconvnet_b200.nn.modules.normalizations import BatchNorm
from warpconvnet_b200.nn.modules.sparse_conv import SparseConv3d


class ConvBlock(nn.Module):
    """conv -> BatchNorm -> (+ residual) -> ReLU; the norm / add / activation tail is ONE fused
    row-streaming pass pair (warpconvnet_b200.nn.modules.BatchNorm, csrc/rownorm.cu)."""

    def __init__(self, cin, cout, kernel_size=3, stride=1, act=True):
        super().__init__()
        self.conv = SparseConv3d(cin, cout, kernel_size, stride, bias=False)
        self.conv.emit_bn_stats = True   # BatchNorm statistics come out of the GEMM epilogue
        self.bn = BatchNorm(cout, relu=act)

    def forward(self, x, residual=None):
        return self.bn(self.conv(x), residual=residual)


class ConvTrBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv_tr = SparseConv3d(cin, cout, 2, 2, transposed=True, bias=False)
        self.conv_tr.emit_bn_stats = True
        self.bn = BatchNorm(cout, relu=True)

    def forward(self, x, target):
        return self.bn(self.conv_tr(x, target))


class BasicBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = ConvBlock(cin, cout, 3)
        self.conv2 = ConvBlock(cout, cout, 3, act=True)   # relu(bn(conv2) + identity)
        self.downsample = ConvBlock(cin, cout, 1, act=False) if cin != cout else None

    def forward(self, x):
        idn = x if self.downsample is None else self.downsample(x)
        return self.conv2(self.conv1(x), residual=idn)


def cat(a: Voxels, b: Voxels) -> Voxels:
    return a.replace(batched_features=torch.cat([a.feature_tensor, b.feature_tensor], dim=1))


class MinkUNet14(nn.Module):
    PLANES = (32, 64, 128, 256, 128, 128, 96, 96)

    def __init__(self, in_channels=3, out_channels=20, init_dim=32):
        super().__init__()
        p = self.PLANES
        self.conv0 = ConvBlock(in_channels, init_dim, 1)
        self.conv1, self.block1 = ConvBlock(init_dim, init_dim, 2, 2), BasicBlock(init_dim, p[0])
        self.conv2, self.block2 = ConvBlock(p[0], p[0], 2, 2), BasicBlock(p[0], p[1])
        self.conv3, self.block3 = ConvBlock(p[1], p[1], 2, 2), BasicBlock(p[1], p[2])
        self.conv4, self.block4 = ConvBlock(p[2], p[2], 2, 2), BasicBlock(p[2], p[3])
        self.convtr4, self.block5 = ConvTrBlock(p[3], p[4]), BasicBlock(p[4] + p[2], p[4])
        self.convtr5, self.block6 = ConvTrBlock(p[4], p[5]), BasicBlock(p[5] + p[1], p[5])
        self.convtr6, self.block7 = ConvTrBlock(p[5], p[6]), BasicBlock(p[6] + p[0], p[6])
        self.convtr7, self.block8 = ConvTrBlock(p[6], p[7]), BasicBlock(p[7] + init_dim, p[7])
        self.final = SparseConv3d(p[7], out_channels, 1, bias=True)

    def forward(self, x: Voxels) -> Voxels:
        p1 = self.conv0(x)
        b1 = self.block1(self.conv1(p1))
        b2 = self.block2(self.conv2(b1))
        b3 = self.block3(self.conv3(b2))
        out = self.block4(self.conv4(b3))
        out = self.block5(cat(self.convtr4(out, b3), b3))
        out = self.block6(cat(self.convtr5(out, b2), b2))
        out = self.block7(cat(self.convtr6(out, b1), b1))
        out = self.block8(cat(self.convtr7(out, p1), p1))
        return self.final(out)


def surface_scene(extent: int, seed: int) -> torch.Tensor:
    """One ScanNet-like synthetic scene: height field over [0, extent)^2."""
    import numpy as np
    rng = np.random.RandomState(seed)
    a, b = rng.uniform(0, 2 * np.pi, size=2)
    u, v = np.meshgrid(np.arange(extent), np.arange(extent), indexing="ij")
    z = np.rint(12 * np.sin(2 * np.pi * u / 180 + a) + 8 * np.cos(2 * np.pi * v / 130 + b)) + 256
    return torch.from_numpy(np.stack([u.reshape(-1), v.reshape(-1), z.reshape(-1)], 1).astype(np.int32))
