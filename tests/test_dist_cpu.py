"""Host-side multi-GPU logic on CPU with the gloo backend, world_size 2 and 3 (no GPU needed): ray
sharding arithmetic, the image gather (equal and ragged blocks) and the flat gradient all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msra_practice_project_b200 import dist as shard


def test_shard_range_partitions():
    for n in (0, 1, 7, 144, 640000, 16777216):
        for world in (1, 2, 3, 4, 8):
            parts = [shard.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    # 800x800 frame: equal blocks of whole pixel rows at 1/2/4/8 GPUs
    for world in (1, 2, 4, 8):
        assert all(shard.shard_range(640000, r, world)[1] == 640000 // world for r in range(world))
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_rays, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_rays * 5, dtype=torch.float32).reshape(n_rays, 5)       # the frame every rank should end with
        b, c = shard.shard_range(n_rays, rank, world)
        out = torch.full((n_rays, 5), -1.0)
        shard.gather_image(full[b:b + c, :3].clone(), full[b:b + c, 3].clone(), full[b:b + c, 4].clone(), out, n_rays, rank, world)
        ok_gather = bool(torch.equal(out, full))
        # in place: the rank's rows already sit in the frame buffer (the composite kernel writes them there): no pack, no staging
        out2 = torch.full((n_rays, 5), -1.0)
        sl = shard.frame_slice(out2, n_rays, rank, world)
        sl.copy_(full[b:b + c])
        shard.gather_image(sl[:, :3], sl[:, 3], sl[:, 4], out2, n_rays, rank, world)
        ok_gather = ok_gather and bool(torch.equal(out2, full))
        # gradient all-reduce of two small "models": every rank holds grad = rank+1 -> mean = (world+1)/2
        torch.manual_seed(0)
        ms = [torch.nn.Linear(3, 4), torch.nn.Linear(4, 2)]
        for m in ms:
            for p in m.parameters():
                p.grad = torch.full_like(p, float(rank + 1))
        shard.allreduce_gradients(ms)
        want = (world + 1) / 2
        ok_grad = all(bool(torch.allclose(p.grad, torch.full_like(p, want))) for m in ms for p in m.parameters())
        q.put((rank, ok_gather, ok_grad))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_rays", [(2, 144), (3, 100), (2, 7)])
def test_gather_and_allreduce_gloo(world, n_rays):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rays, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(g and a for _, g, a in res), res
