# SPDX-License-Identifier: Apache-2.0
"""world_size-2 gloo test of the scene-sharded data-parallel logic: each rank
computes the weight gradient of its own scenes (CPU oracle stands in for the device kernels — this
test covers the host-side sharding + collective, not the kernels), the flat bucket all-reduce
must reproduce the single-process gradient over the whole batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene(seed, n=300, side=10):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(side ** 3, generator=g)[:n].numpy()
    c = np.stack([idx // (side * side), (idx // side) % side, idx % side], 1).astype(np.int32)
    x = torch.randn(n, 4, generator=g).double().numpy()
    gy = torch.randn(n, 8, generator=g).double().numpy()
    return c, x, gy


def _wgrad(scenes):
    from oracle import conv as oconv
    from oracle import kernel_map as okm
    bc = okm.batch_indexed([s[0] for s in scenes])
    km = okm.generate_kernel_map(bc, bc, (1, 1, 1), (3, 3, 3))
    x = np.concatenate([s[1] for s in scenes])
    gy = np.concatenate([s[2] for s in scenes])
    w = np.zeros((27, 4, 8))
    _, dw = oconv.backward(gy, x, w, km["in_maps"], km["out_maps"], km["offsets"])
    return dw, len(bc)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from warpconvnet_b200.dist import FlatGradBucket, all_reduce_wgrad, shard_scenes
    mine = shard_scenes(4, rank, world)
    dw, n_vox = _wgrad([_scene(s) for s in mine])
    w = torch.nn.Parameter(torch.zeros(27, 4, 8))
    b = torch.nn.Parameter(torch.zeros(8))
    bucket = FlatGradBucket([w, b])
    w.grad.copy_(dw.float())
    b.grad.fill_(float(rank + 1))
    bucket.all_reduce()
    t = dw.float().clone()
    all_reduce_wgrad([t])
    count = torch.tensor([n_vox], dtype=torch.int64)
    dist.all_reduce(count)
    if rank == 0:
        q.put((w.grad.clone().numpy(), b.grad.clone().numpy(), t.numpy(), int(count)))
    dist.barrier()
    dist.destroy_process_group()


def test_scene_sharding_covers_batch():
    sys.path.insert(0, ROOT)
    from warpconvnet_b200.dist import shard_scenes
    for world in (1, 2, 4, 8):
        got = sorted(s for r in range(world) for s in shard_scenes(8, r, world))
        assert got == list(range(8))
    with pytest.raises(ValueError):
        shard_scenes(8, 2, 2)


def test_two_rank_wgrad_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    wgrad, bgrad, t, count = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full, n_full = _wgrad([_scene(s) for s in range(4)])
    assert count == n_full
    assert np.allclose(wgrad, full.float().numpy(), rtol=1e-5, atol=1e-5)
    assert np.allclose(t, full.float().numpy(), rtol=1e-5, atol=1e-5)
    assert np.allclose(bgrad, 3.0)


def test_flat_grad_bucket_peer_request_without_a_process_group_uses_the_plain_buffer():
    """peer=True needs an initialised multi-rank group (symmetric memory); in a single process the
    bucket silently keeps its ordinary buffer and all_reduce is a no-op."""
    import torch
    from warpconvnet_b200.dist import FlatGradBucket
    lin = torch.nn.Linear(4, 3)
    bucket = FlatGradBucket(lin.parameters(), peer=True)
    assert bucket.peer is None and bucket.flat.numel() == 4 * 3 + 3
    lin(torch.ones(2, 4)).sum().backward()
    assert lin.weight.grad.data_ptr() == bucket.flat.data_ptr()
    assert bucket.all_reduce(average=True) is None
