# SPDX-License-Identifier: Apache-2.0
"""Generates tests/golden/pointconv_*.npz from the REFERENCE's own PointConv run on CPU
(upstream pins PointConv only by shape; this fixture pins its VALUES).

Runs only where /root/reference is mounted. The reference package is imported with a stub for its
compiled extension (never reached: kNN on CPU is cdist + topk, the MLPs are torch) and with
baseline/torch_scatter_shim.py standing in for torch_scatter.segment_csr (plain torch
scatter_reduce). Saved per case: input coordinates / features / offsets, the module's state_dict,
its constructor arguments, the output features and the kNN indices.
    python tests/golden/make_golden_pointconv.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_adapter  # noqa: E402

spec = importlib.util.spec_from_file_location("torch_scatter",
                                              os.path.join(ROOT, "baseline", "torch_scatter_shim.py"))
shim = importlib.util.module_from_spec(spec)
spec.loader.exec_module(shim)
sys.modules["torch_scatter"] = shim
ref_adapter.load()  # stubs warpconvnet._C, puts /root/reference on sys.path

from warpconvnet.geometry.coords.search.search_configs import RealSearchConfig  # noqa: E402
from warpconvnet.geometry.types.points import Points  # noqa: E402
from warpconvnet.nn.modules.point_conv import PointConv  # noqa: E402

CASES = {
    "pointconv_knn8": dict(cin=16, cout=32, k=8, sizes=(180, 140), kw={}),
    "pointconv_knn16_relpos": dict(cin=8, cout=24, k=16, sizes=(300,),
                                   kw=dict(use_rel_pos=True, reductions=("mean", "max"))),
}

for name, cfg in CASES.items():
    torch.manual_seed(7)
    g = torch.Generator().manual_seed(11)
    coords = [torch.rand(n, 3, generator=g) for n in cfg["sizes"]]
    feats = [torch.randn(n, cfg["cin"], generator=g) for n in cfg["sizes"]]
    pc = Points(coords, feats)
    conv = PointConv(cfg["cin"], cfg["cout"], RealSearchConfig("knn", knn_k=cfg["k"]), **cfg["kw"])
    conv.eval()
    with torch.no_grad():
        out = conv(pc)
    nbrs = pc.neighbors(RealSearchConfig("knn", knn_k=cfg["k"]))
    state = {k: v.detach().cpu().numpy() for k, v in conv.state_dict().items()}
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        coords=torch.cat(coords).numpy(), feats=torch.cat(feats).numpy(),
        offsets=np.cumsum([0] + list(cfg["sizes"])).astype(np.int64),
        out=out.feature_tensor.detach().numpy(),
        knn=nbrs.neighbor_indices.reshape(-1, cfg["k"]).numpy(),
        state_keys=np.array(list(state.keys())), **{"p__" + k: v for k, v in state.items()})
    print(name, "out", tuple(out.feature_tensor.shape), "params", list(state.keys()))
