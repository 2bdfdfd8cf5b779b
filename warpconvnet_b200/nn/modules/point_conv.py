# SPDX-License-Identifier: Apache-2.0
"""``PointConv`` (drop-in for warpconvnet/nn/modules/point_conv.py:36-282): neighbour search on
the device grid-kNN kernel -> gather edge features -> edge MLP -> row reduction -> output MLP.
Same constructor arguments, sub-module names (``edge_transform_mlp``, ``out_transform_mlp``) and
``forward(in_pc, query_pc=None) -> Points``."""
from __future__ import annotations

import warnings
from typing import List, Literal, Optional

import torch
import torch.nn as nn

from warpconvnet_b200.geometry.base.batched import Coords
from warpconvnet_b200.geometry.coords.search.search_configs import (RealSearchConfig,
                                                                    RealSearchMode)
from warpconvnet_b200.geometry.types.points import Points
from warpconvnet_b200.nn.encodings import SinusoidalEncoding
from warpconvnet_b200.nn.modules.base_module import BaseSpatialModule
from warpconvnet_b200.nn.modules.mlp import MLPBlock
from warpconvnet_b200.ops.reductions import REDUCTION_TYPES_STR, REDUCTIONS, row_reduction


def _get_module_input_channel(module: nn.Module) -> int:
    if isinstance(module, nn.Linear):
        return module.in_features
    if isinstance(module, nn.Sequential):
        return _get_module_input_channel(module[0])
    for _, child in module.named_children():
        return _get_module_input_channel(child)
    raise ValueError(f"Unsupported module type: {type(module)}")


class PointConv(BaseSpatialModule):
    def __init__(self, in_channels: int, out_channels: int,
                 neighbor_search_args: RealSearchConfig,
                 pooling_reduction: Optional[REDUCTIONS] = None,
                 pooling_voxel_size: Optional[float] = None,
                 edge_transform_mlp: Optional[nn.Module] = None,
                 out_transform_mlp: Optional[nn.Module] = None, mlp_block=MLPBlock,
                 hidden_dim: Optional[int] = None, channel_multiplier: int = 2,
                 use_rel_pos: bool = False, use_rel_pos_encode: bool = False,
                 pos_encode_dim: int = 32, pos_encode_range: float = 4,
                 reductions: List[REDUCTION_TYPES_STR] = ("mean",),
                 out_point_type: Literal["provided", "downsample", "same"] = "same",
                 provided_in_channels: Optional[int] = None, bias: bool = True):
        super().__init__()
        assert isinstance(reductions, (tuple, list)) and len(reductions) > 0, \
            f"reductions must be a list or tuple of length > 0, got {reductions}"
        if out_point_type == "provided":
            assert pooling_reduction is None and pooling_voxel_size is None
            assert provided_in_channels is not None, \
                "provided_in_channels must be provided for provided type"
        elif out_point_type == "downsample":
            assert pooling_reduction is not None and pooling_voxel_size is not None, \
                "pooling_reduction and pooling_voxel_size must be provided for downsample type"
            assert provided_in_channels is None
            if (neighbor_search_args.mode == RealSearchMode.RADIUS
                    and neighbor_search_args.radius < pooling_voxel_size * (3 ** 0.5)):
                warnings.warn(f"neighbor search radius {neighbor_search_args.radius} is less than "
                              f"sqrt(3) times the downsample voxel size {pooling_voxel_size}",
                              stacklevel=2)
        elif out_point_type == "same":
            assert pooling_reduction is None and pooling_voxel_size is None
            assert provided_in_channels is None
        assert isinstance(neighbor_search_args, RealSearchConfig)
        self.reductions = reductions
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.use_rel_pos = use_rel_pos
        self.use_rel_pos_encode = use_rel_pos_encode
        self.out_point_feature_type = out_point_type
        self.neighbor_search_args = neighbor_search_args
        self.pooling_reduction = pooling_reduction
        self.pooling_voxel_size = pooling_voxel_size
        # concat_input=False so the encoding width is pos_encode_dim * 3, the number the
        # reference's own channel bookkeeping uses (point_conv.py:160-164,213-219)
        self.positional_encoding = SinusoidalEncoding(pos_encode_dim, data_range=pos_encode_range,
                                                      concat_input=False)
        if provided_in_channels is None:
            provided_in_channels = in_channels
        if hidden_dim is None:
            hidden_dim = channel_multiplier * max(out_channels, in_channels)
        if edge_transform_mlp is None:
            edge_in = in_channels + provided_in_channels
            if use_rel_pos_encode:
                edge_in += pos_encode_dim * 3
            elif use_rel_pos:
                edge_in += 3
            edge_transform_mlp = mlp_block(in_channels=edge_in, out_channels=out_channels,
                                           hidden_channels=hidden_dim, bias=bias)
        self.edge_transform_mlp = edge_transform_mlp
        self.edge_mlp_in_channels = _get_module_input_channel(edge_transform_mlp)
        if out_transform_mlp is None:
            out_transform_mlp = mlp_block(in_channels=out_channels * len(reductions),
                                          out_channels=out_channels, hidden_channels=hidden_dim,
                                          bias=bias)
        self.out_transform_mlp = out_transform_mlp

    def __repr__(self):
        s = f"{self.__class__.__name__}(in_channels={self.in_channels} out_channels={self.out_channels}"
        if self.use_rel_pos_encode:
            s += f" rel_pos_encode={self.use_rel_pos_encode}"
        if self.pooling_reduction is not None:
            s += f" pooling={self.pooling_reduction}"
        return s + f" neighbor={self.neighbor_search_args})"

    def forward(self, in_pc: Points, query_pc: Optional[Points] = None) -> Points:
        if self.out_point_feature_type == "provided":
            assert query_pc is not None, \
                "query_point_features must be provided for the provided type"
        elif self.out_point_feature_type == "downsample":
            raise NotImplementedError(
                "out_point_type='downsample' needs Points.voxel_downsample, which is outside the "
                "hot path (SURVEY.md §2a)")
        else:
            assert query_pc is None
            query_pc = in_pc
        in_c, q_c = in_pc.num_channels, query_pc.num_channels
        expect = (in_c + q_c + self.use_rel_pos_encode * self.positional_encoding.num_channels * 3
                  + (not self.use_rel_pos_encode) * self.use_rel_pos * 3)
        assert expect == self.edge_mlp_in_channels, \
            (f"input features {tuple(in_pc.feature_tensor.shape)} and query features "
             f"{tuple(query_pc.feature_tensor.shape)} do not match the edge_transform_mlp input "
             f"channels {self.edge_mlp_in_channels}")

        neighbors = in_pc.neighbors(query_coords=query_pc.batched_coordinates,
                                    search_args=self.neighbor_search_args)
        idx = neighbors.neighbor_indices.long().view(-1)
        row_splits = neighbors.neighbor_row_splits
        num_reps = row_splits[1:] - row_splits[:-1]

        rep_in = in_pc.feature_tensor[idx]
        self_feats = torch.repeat_interleave(
            query_pc.feature_tensor.view(-1, q_c).contiguous(), num_reps, dim=0)
        edge = [rep_in, self_feats]
        if self.use_rel_pos or self.use_rel_pos_encode:
            rel = (in_pc.coordinate_tensor.view(-1, 3)[idx]
                   - torch.repeat_interleave(query_pc.coordinate_tensor.view(-1, 3).contiguous(),
                                             num_reps, dim=0))
            edge.append(self.positional_encoding(rel).to(rep_in.dtype)
                        if self.use_rel_pos_encode else rel.to(rep_in.dtype))
        edge = self.edge_transform_mlp(torch.cat(edge, dim=1))
        out = torch.cat([row_reduction(edge, row_splits, reduction=r) for r in self.reductions],
                        dim=-1)
        out = self.out_transform_mlp(out)
        return Points(batched_coordinates=Coords(query_pc.coordinate_tensor, query_pc.offsets),
                      batched_features=out, **query_pc.extra_attributes)
