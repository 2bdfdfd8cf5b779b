# SPDX-License-Identifier: Apache-2.0
"""torchrun --nproc-per-node N tools/exp_peer_allreduce.py : wcn_peer_allreduce_f32 against NCCL —
results (bit-level agreement is not expected: the summation order differs), repeated calls,
graph replay, and device time of both for the dW of one 27 x 128 x 128 layer."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from warpconvnet_b200.dist import PeerAllReduce  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    for numel in (27 * 128 * 128, 1000, 4 * 1024 * 1024 + 4):
        par = PeerAllReduce(numel, dev)
        worst = 0.0
        for it in range(5):
            g = torch.Generator(device=dev).manual_seed(1000 * it + rank)
            src = torch.randn(numel, device=dev, generator=g)
            want = src.clone()
            dist.all_reduce(want)
            par.buffer.copy_(src)
            got = par.all_reduce_()
            worst = max(worst, float((got - want).abs().max()))
        # integers: exact whatever the order
        par.buffer.copy_(torch.arange(numel, device=dev, dtype=torch.float32) % 1024 + rank)
        got = par.all_reduce_()
        want = (torch.arange(numel, device=dev, dtype=torch.float32) % 1024) * world \
            + world * (world - 1) / 2
        exact = bool(torch.equal(got, want))
        out[f"n{numel}"] = {"max_abs_vs_nccl": worst, "exact_on_integers": exact}
        assert exact and worst < 1e-4, out

    numel = 27 * 128 * 128
    par = PeerAllReduce(numel, dev)
    ref = torch.zeros(numel, device=dev)

    # graph capture + replay (barrier epochs live on the device)
    par.buffer.fill_(1.0)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.graph(graph):
        par.all_reduce_()
    for i in range(3):
        par.buffer.fill_(float(rank + i))
        graph.replay()
        torch.cuda.synchronize()
        want = sum(r + i for r in range(world))
        assert bool((par.buffer == want).all()), (i, par.buffer[:4], want)
    out["graph_replay"] = "ok"

    def timeit(fn, reps=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / reps * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(float(t), 2)

    out["us_peer_1.77MB"] = timeit(lambda: par.all_reduce_())
    out["us_nccl_1.77MB"] = timeit(lambda: dist.all_reduce(ref))
    for ctas in (16, 64, 128):
        par.n_ctas = ctas
        out[f"us_peer_{ctas}ctas"] = timeit(lambda: par.all_reduce_())
    big = PeerAllReduce(8 * 1024 * 1024, dev, n_ctas=64)
    refb = torch.zeros(8 * 1024 * 1024, device=dev)
    out["us_peer_32MB_64ctas"] = timeit(lambda: big.all_reduce_(), 50)
    out["us_nccl_32MB"] = timeit(lambda: dist.all_reduce(refb), 50)
    if rank == 0:
        print(json.dumps(out), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
