# SPDX-License-Identifier: Apache-2.0
"""Where does the UNMODIFIED reference crash on this B200? One C3-S fwd+bwd through its public API
in its default ``auto`` mode with faulthandler and its auto-tune log switched on (the Python stack
at the SIGSEGV names the backend and tile).  python -X faulthandler tools/ref_diag.py [mode]"""
import faulthandler
import os
import sys

faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("WARPCONVNET_BENCHMARK_CACHE_DIR", os.path.join(ROOT, "gpurun_out", "ref_cache_diag"))
os.environ["WARPCONVNET_AUTOTUNE_LOG"] = "true"
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
import warpconvnet  # noqa: E402,F401
from warpconvnet.geometry.types.voxels import Voxels  # noqa: E402
from warpconvnet.nn.modules.sparse_conv import SparseConv3d  # noqa: E402
from ref_gpu_bench import surface  # noqa: E402

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 448
c = surface(n_side, 0).cuda()
n = len(c)
x = torch.randn(n, 128, device="cuda").bfloat16()
conv = SparseConv3d(128, 128, 3, bias=False).cuda()
vox = Voxels([c], [x])
gy = torch.randn(n, 128, device="cuda").bfloat16()
for it in range(3):
    conv.weight.grad = None
    v = vox.replace(batched_features=vox.feature_tensor.detach().requires_grad_(True))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = conv(v)
    print("forward ok", it, flush=True)
    y.feature_tensor.backward(gy)
    torch.cuda.synchronize()
    print("backward ok", it, flush=True)
print("DONE")
