# SPDX-License-Identifier: Apache-2.0
"""torchrun --nproc-per-node N tools/exp_peer_step.py : where the time of the dW all-reduce goes
inside the C3-S step (wgrad -> all-reduce || dgrad), graph replays, max over ranks."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from warpconvnet_b200 import _ops  # noqa: E402
from warpconvnet_b200.dist import PeerAllReduce  # noqa: E402
from warpconvnet_b200.geometry.coords.search.torch_discrete import generate_kernel_map  # noqa: E402
from warpconvnet_b200.nn.functional.sparse_conv import sparse_conv_wgrad  # noqa: E402

K, CIN, COUT, KS = bench.K, bench.CIN, bench.COUT, bench.KS


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    import numpy as np
    coords = bench.make_coords("S", seed=rank)
    n = len(coords)
    x_h, w_h, gy_h = bench.make_tensors(n, seed=rank)
    bc = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), coords], 1)).to(dev)
    x, w, gy = x_h.to(dev).bfloat16(), w_h.to(dev).bfloat16(), gy_h.to(dev).bfloat16()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    km = generate_kernel_map(bc, bc, (1, 1, 1), (KS,) * 3, same_coords=True)
    plan = km.fwd_plan(n)
    img, img_t = _ops.weight_image_pair(w.view(K, 1, CIN, COUT), K, 1, CIN, COUT, w.dtype)
    bplan, kflip = km.bwd_plan(n)
    par = PeerAllReduce(K * CIN * COUT, dev, n_ctas=int(os.environ.get("AR_CTAS", 16)))
    side = torch.cuda.Stream(device=dev)
    plain = torch.zeros(K * CIN * COUT, device=dev)

    def wgrad(buf):
        return sparse_conv_wgrad(x, gy, (K, CIN, COUT), km, out=buf)

    def dgrad():
        return _ops.gather_gemm(gy, img_t, bplan, 1, COUT, CIN, kflip=kflip)

    def v_wgrad_plain():
        wgrad(plain)

    def v_wgrad_symm():
        wgrad(par.buffer)

    def v_dgrad():
        dgrad()

    def v_ar():
        par.all_reduce_()

    def v_nccl():
        dist.all_reduce(plain)

    def v_wd_none():
        wgrad(par.buffer)
        dgrad()

    def v_wd_serial():
        wgrad(par.buffer)
        par.all_reduce_()
        dgrad()

    def v_wd_after():
        wgrad(par.buffer)
        dgrad()
        par.all_reduce_()

    def v_wd_side():
        wgrad(par.buffer)
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            par.all_reduce_()
        dgrad()
        cur.wait_stream(side)

    def v_wd_side_dfirst():
        wgrad(par.buffer)
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        dgrad()
        with torch.cuda.stream(side):
            par.all_reduce_()
        cur.wait_stream(side)

    def v_wd_nccl():
        wgrad(plain)
        work = dist.all_reduce(plain, async_op=True)
        dgrad()
        work.wait()

    def v_dw_side():   # dgrad first, then wgrad; all-reduce at the end (nothing to hide under)
        dgrad()
        wgrad(par.buffer)
        par.all_reduce_()

    if "--trace" in sys.argv:
        # kernel timeline (CUPTI through torch.profiler) of graph replays of the shipped order:
        # is the all-reduce kernel inside the dgrad kernel's interval?
        from torch.profiler import ProfilerActivity, profile
        for _ in range(3):
            v_wd_side_dfirst()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="relaxed"):
            v_wd_side_dfirst()
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(4):
                flush.fill_(1)
                g.replay()
            torch.cuda.synchronize()
        rows = []
        for ev in prof.events():
            if ev.device_type is not None and "CUDA" in str(ev.device_type):
                rows.append((ev.time_range.start, ev.time_range.end - ev.time_range.start, ev.name[:60]))
        rows.sort()
        t0 = rows[0][0]
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"peer_timeline_rank{rank}_of{world}.txt"), "w") as f:
            f.write("start_us  dur_us  kernel\n")
            for st, du, nm in rows:
                f.write(f"{st - t0:9.1f} {du:7.1f}  {nm}\n")
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)

    out = {}
    for name, fn in list(locals().items()):
        if not name.startswith("v_"):
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="relaxed"):
            fn()
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        evs = []
        for _ in range(30):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            g.replay()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        t = torch.tensor([sum(s.elapsed_time(e) for s, e in evs) / len(evs) * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name[2:]] = round(float(t), 1)
        del g
    if rank == 0:
        print(json.dumps(out), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
