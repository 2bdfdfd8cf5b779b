# SPDX-License-Identifier: Apache-2.0
"""Which of the reference's own backends survive on this B200? Runs the UNMODIFIED reference
(baseline/_ref) once per (direction, algorithm) in a fresh subprocess — selected through the
reference's own environment switches WARPCONVNET_{FWD,DGRAD,WGRAD}_ALGO_MODE (constants.py:136-160)
— on the C3-S layer and reports ok / crash / error + the fwd+bwd time. The auto-tuner of the
reference segfaulted inside its backward candidate sweep on the first head-to-head attempt
(gpurun_out/r2a_ref.log); this probe finds the candidates to exclude so that the comparison can
run with the rest of its pool.   python tools/ref_probe_algos.py > gpurun_out/ref_probe.jsonl"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, time
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import warpconvnet
from warpconvnet.geometry.types.voxels import Voxels
from warpconvnet.nn.modules.sparse_conv import SparseConv3d
from ref_gpu_bench import surface
c = surface(448, 0).cuda(); n = len(c)
x = torch.randn(n, 128, device="cuda").bfloat16()
kw = {}
if os.environ.get("PROBE_DGRAD"): kw["dgrad_algo"] = [os.environ["PROBE_DGRAD"]]
if os.environ.get("PROBE_WGRAD"): kw["wgrad_algo"] = [os.environ["PROBE_WGRAD"]]
conv = SparseConv3d(128, 128, 3, bias=False, **kw).cuda()
vox = Voxels([c], [x])
gy = torch.randn(n, 128, device="cuda").bfloat16()
def step():
    conv.weight.grad = None
    v = vox.replace(batched_features=vox.feature_tensor.detach().requires_grad_(True))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = conv(v)
    y.feature_tensor.backward(gy)
for _ in range(4): step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): step()
e.record(); torch.cuda.synchronize()
print("RESULT", json.dumps({"fwd_bwd_ms": s.elapsed_time(e) / 10}))
'''
ALGOS = ["mask_gemm", "mask_gemm_fwd_as_dgrad", "cutlass_implicit_gemm", "cute_grouped",
         "cutlass_grouped_hybrid", "explicit_gemm", "implicit_gemm"]


def run(env_extra, timeout=240):
    env = dict(os.environ, WARPCONVNET_AUTOTUNE_LOG="false",
               WARPCONVNET_BENCHMARK_CACHE_DIR=os.path.join(ROOT, "gpurun_out", "ref_cache_probe"),
               **env_extra)
    try:
        p = subprocess.run([sys.executable, "-c", f"ROOT={ROOT!r}\n" + CHILD], env=env,
                           capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {"status": "timeout"}
    for line in p.stdout.splitlines():
        if line.startswith("RESULT"):
            return dict(json.loads(line[7:]), status="ok")
    tail = (p.stderr or "").strip().splitlines()[-1:] or [""]
    return {"status": "crash" if p.returncode < 0 else "error", "rc": p.returncode,
            "last_stderr_line": tail[0][:200]}


def main():
    """forward stays on the reference's default auto pool (7 candidates, runs clean); one backward
    direction at a time is pinned to one backend family through the constructor (a list of names
    is the reference's documented way to restrict the pool, nn/modules/sparse_conv.py:130-137),
    the other direction to explicit_gemm."""
    which = sys.argv[1:] or ["DGRAD", "WGRAD"]
    for direction in which:
        other = "WGRAD" if direction == "DGRAD" else "DGRAD"
        for algo in ALGOS:
            if direction == "WGRAD" and algo == "mask_gemm_fwd_as_dgrad":
                continue
            r = run({f"PROBE_{direction}": algo, f"PROBE_{other}": "explicit_gemm"})
            r.update(direction=direction.lower(), algo=algo)
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
