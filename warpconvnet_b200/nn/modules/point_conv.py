# SPDX-License-Identifier: Apache-2.0
"""Continuous convolution on point clouds.

API-compatible with the reference's ``PointConv`` (warpconvnet/nn/modules/point_conv.py:36-282):
identical constructor keywords, sub-module names ``edge_transform_mlp`` / ``out_transform_mlp`` and
``forward(in_pc, query_pc=None) -> Points``. Pipeline: neighbour search on the device grid
(kNN or radius, ``csrc/knn.cu``) -> per-edge features [neighbour | query | relative position] ->
edge MLP -> one reduction per requested type over each query's neighbours -> output MLP.
"""
from __future__ import annotations

import warnings
from typing import Optional, Sequence

import torch
import torch.nn as nn

from warpconvnet_b200.geometry.base.batched import Coords
from warpconvnet_b200.geometry.coords.search import search_configs as _cfg
from warpconvnet_b200.geometry.types.points import Points
from warpconvnet_b200.nn.encodings import SinusoidalEncoding
from warpconvnet_b200.nn.modules.base_module import BaseSpatialModule
from warpconvnet_b200.nn.modules.mlp import MLPBlock
from warpconvnet_b200.ops import reductions as _red

RealSearchConfig = _cfg.RealSearchConfig
RealSearchMode = _cfg.RealSearchMode
REDUCTIONS = _red.REDUCTIONS
REDUCTION_TYPES_STR = _red.REDUCTION_TYPES_STR
row_reduction = _red.row_reduction


def _first_linear_width(module: nn.Module) -> int:
    """Input width of the first ``nn.Linear`` reachable through first children."""
    node = module
    while not isinstance(node, nn.Linear):
        children = list(node.children())
        if not children:
            raise ValueError(f"Unsupported module type: {type(module)}")
        node = children[0]
    return node.in_features


def _check_output_mode(mode, search, pooling_reduction, pooling_voxel_size, provided_in_channels):
    pooled = pooling_reduction is not None and pooling_voxel_size is not None
    unpooled = pooling_reduction is None and pooling_voxel_size is None
    if mode == "provided":
        assert unpooled
        assert provided_in_channels is not None, \
            "provided_in_channels must be provided for provided type"
    elif mode == "downsample":
        assert pooled, \
            "pooling_reduction and pooling_voxel_size must be provided for downsample type"
        assert provided_in_channels is None
        too_small = (search.mode == RealSearchMode.RADIUS
                     and search.radius < pooling_voxel_size * 3 ** 0.5)
        if too_small:
            warnings.warn(f"neighbor search radius {search.radius} is less than sqrt(3) times "
                          f"the downsample voxel size {pooling_voxel_size}", stacklevel=3)
    elif mode == "same":
        assert unpooled and provided_in_channels is None


class PointConv(BaseSpatialModule):
    def __init__(self, in_channels: int, out_channels: int,
                 neighbor_search_args: RealSearchConfig, pooling_reduction=None,
                 pooling_voxel_size: Optional[float] = None,
                 edge_transform_mlp: Optional[nn.Module] = None,
                 out_transform_mlp: Optional[nn.Module] = None, mlp_block=MLPBlock,
                 hidden_dim: Optional[int] = None, channel_multiplier: int = 2,
                 use_rel_pos: bool = False, use_rel_pos_encode: bool = False,
                 pos_encode_dim: int = 32, pos_encode_range: float = 4,
                 reductions: Sequence[str] = ("mean",), out_point_type: str = "same",
                 provided_in_channels: Optional[int] = None, bias: bool = True):
        super().__init__()
        assert isinstance(reductions, (tuple, list)) and len(reductions) > 0, \
            f"reductions must be a list or tuple of length > 0, got {reductions}"
        assert isinstance(neighbor_search_args, RealSearchConfig)
        _check_output_mode(out_point_type, neighbor_search_args, pooling_reduction,
                           pooling_voxel_size, provided_in_channels)
        settings = dict(reductions=reductions, in_channels=in_channels, out_channels=out_channels,
                        use_rel_pos=use_rel_pos, use_rel_pos_encode=use_rel_pos_encode,
                        out_point_feature_type=out_point_type,
                        neighbor_search_args=neighbor_search_args,
                        pooling_reduction=pooling_reduction, pooling_voxel_size=pooling_voxel_size)
        for key, value in settings.items():
            setattr(self, key, value)
        # encoding width = 3 * pos_encode_dim (concat_input=False), the number the reference's
        # channel bookkeeping assumes (point_conv.py:160-164,213-219)
        self.positional_encoding = SinusoidalEncoding(pos_encode_dim, data_range=pos_encode_range,
                                                      concat_input=False)
        query_width = in_channels if provided_in_channels is None else provided_in_channels
        width = hidden_dim or channel_multiplier * max(out_channels, in_channels)
        position_width = 3 * pos_encode_dim if use_rel_pos_encode else (3 if use_rel_pos else 0)
        if edge_transform_mlp is None:
            edge_transform_mlp = mlp_block(in_channels=in_channels + query_width + position_width,
                                           out_channels=out_channels, hidden_channels=width,
                                           bias=bias)
        self.edge_transform_mlp = edge_transform_mlp
        self.edge_mlp_in_channels = _first_linear_width(self.edge_transform_mlp)
        if out_transform_mlp is None:
            out_transform_mlp = mlp_block(in_channels=out_channels * len(reductions),
                                          out_channels=out_channels, hidden_channels=width,
                                          bias=bias)
        self.out_transform_mlp = out_transform_mlp

    def __repr__(self):
        extras = []
        if self.use_rel_pos_encode:
            extras.append(f"rel_pos_encode={self.use_rel_pos_encode}")
        if self.pooling_reduction is not None:
            extras.append(f"pooling={self.pooling_reduction}")
        extras.append(f"neighbor={self.neighbor_search_args}")
        return (f"{type(self).__name__}(in_channels={self.in_channels} "
                f"out_channels={self.out_channels} " + " ".join(extras) + ")")

    def _position_width(self) -> int:
        if self.use_rel_pos_encode:
            return 3 * self.positional_encoding.num_channels
        return 3 if self.use_rel_pos else 0

    def forward(self, in_pc: Points, query_pc: Optional[Points] = None) -> Points:
        mode = self.out_point_feature_type
        if mode == "downsample":
            raise NotImplementedError(
                "out_point_type='downsample' needs Points.voxel_downsample, which is outside the "
                "hot path (SURVEY.md §2a)")
        if mode == "provided":
            assert query_pc is not None, \
                "query_point_features must be provided for the provided type"
        else:
            assert query_pc is None
            query_pc = in_pc
        q_width = query_pc.num_channels
        expected = in_pc.num_channels + q_width + self._position_width()
        assert expected == self.edge_mlp_in_channels, \
            (f"input features {tuple(in_pc.feature_tensor.shape)} and query features "
             f"{tuple(query_pc.feature_tensor.shape)} do not match the edge_transform_mlp input "
             f"channels {self.edge_mlp_in_channels}")

        found = in_pc.neighbors(query_coords=query_pc.batched_coordinates,
                                search_args=self.neighbor_search_args)
        nbr = found.neighbor_indices.long().view(-1)
        splits = found.neighbor_row_splits
        per_query = splits[1:] - splits[:-1]

        def per_edge(t):  # one row per (query, neighbour) edge, query value repeated
            return torch.repeat_interleave(t.contiguous(), per_query, dim=0)

        columns = [in_pc.feature_tensor[nbr], per_edge(query_pc.feature_tensor.view(-1, q_width))]
        if self.use_rel_pos or self.use_rel_pos_encode:
            offset = (in_pc.coordinate_tensor.view(-1, 3)[nbr]
                      - per_edge(query_pc.coordinate_tensor.view(-1, 3)))
            if self.use_rel_pos_encode:
                offset = self.positional_encoding(offset)
            columns.append(offset.to(columns[0].dtype))
        edges = self.edge_transform_mlp(torch.cat(columns, dim=1))
        pooled = [row_reduction(edges, splits, reduction=kind) for kind in self.reductions]
        out = self.out_transform_mlp(torch.cat(pooled, dim=-1))
        return Points(batched_coordinates=Coords(query_pc.coordinate_tensor, query_pc.offsets),
                      batched_features=out, **query_pc.extra_attributes)
