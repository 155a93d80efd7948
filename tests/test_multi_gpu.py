"""Multi-GPU (NCCL) checks of SURVEY 8e on real devices: needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`).
  * G-invariance of the sharded render: every rank renders its pixel rows, the gathered frame is bit-identical to the
    single-GPU frame (rays never interact; every rank slices the same global jitter).
  * G-invariance of the fused training step: 2 ranks x half a batch with one all-reduce of the flat gradient bucket follow
    the same loss trajectory / weights as 1 rank x the whole batch (up to the summation order of the gradients).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _setup():
    from msra_practice_project_b200 import models, pigan_render
    torch.manual_seed(0)
    c, f = models.damp_nerf_(models.NeRF()).cuda(), models.damp_nerf_(models.NeRF()).cuda()
    pose = pigan_render.camera_pos_to_transform_matrix(4.0, 0.3, -0.5)
    return c, f, pose


def _single(q):
    from msra_practice_project_b200 import dist as shard, nerf_render, ops
    from msra_practice_project_b200.train_step import NerfTrainStep
    torch.cuda.set_device(0)
    c, f, pose = _setup()
    w, h, sc, sf = 48, 40, 16, 24
    g = torch.Generator().manual_seed(3)
    t_full = torch.rand(w * h, sc, generator=g).cuda()
    with torch.no_grad():
        o = nerf_render.render_image_device(w, h, w * 1.3875, pose, 2.0, 6.0, c, f, sc, sf, t_rand=t_full)
    frame = torch.cat([o[3], o[4][:, None], o[5][:, None]], -1).cpu()
    nb = 512
    rays = ops.raygen(w, h, w * 1.3875, pose, 100, nb)
    target = (torch.rand(nb, 3, generator=g) * 0.5 + 0.25).cuda()
    ts = [torch.rand(nb, sc, generator=g).cuda() for _ in range(4)]
    step = NerfTrainStep(c, f, 2.0, 6.0, sc, sf, nb, learning_rate=5e-4, graph=True)
    losses = [float(step(rays, target, t_rand=t)[0]) for t in ts]
    q.put(("single", frame.numpy(), losses, step.params.cpu().numpy()))


def _rank(rank, world, port, q):
    from msra_practice_project_b200 import dist as shard, ops
    from msra_practice_project_b200.train_step import NerfTrainStep
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        c, f, pose = _setup()
        w, h, sc, sf = 48, 40, 16, 24
        g = torch.Generator().manual_seed(3)
        t_full = torch.rand(w * h, sc, generator=g).cuda()
        rgb, depth, acc = shard.render_image_sharded(w, h, w * 1.3875, pose, 2.0, 6.0, c, f, sc, sf, t_rand_full=t_full)
        frame = torch.cat([rgb.reshape(-1, 3), depth.reshape(-1, 1), acc.reshape(-1, 1)], -1).cpu()
        nb = 512
        b0, cnt = shard.shard_range(nb, rank, world)
        rays = ops.raygen(w, h, w * 1.3875, pose, 100, nb)[b0:b0 + cnt].contiguous()
        target = (torch.rand(nb, 3, generator=g) * 0.5 + 0.25).cuda()[b0:b0 + cnt].contiguous()
        ts = [torch.rand(nb, sc, generator=g).cuda()[b0:b0 + cnt].contiguous() for _ in range(4)]
        step = NerfTrainStep(c, f, 2.0, 6.0, sc, sf, cnt, learning_rate=5e-4, graph=True)
        losses = []
        for t in ts:
            l = step(rays, target, t_rand=t)[0].clone()
            dist.all_reduce(l)                                   # each rank reports its share of the global-batch loss
            losses.append(float(l))
        if rank == 0:
            q.put(("multi", frame.numpy(), losses, step.params.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_render_and_training_match_single_gpu():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p0 = ctx.Process(target=_single, args=(q,))
    p0.start()
    single = q.get(timeout=300)
    p0.join(timeout=60)
    port = _free_port()
    procs = [ctx.Process(target=_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    multi = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert single[0] == "single" and multi[0] == "multi"
    assert np.array_equal(single[1], multi[1]), "sharded frame differs from the single-GPU frame"
    print("losses 1 GPU", single[2], "2 GPUs", multi[2])
    np.testing.assert_allclose(multi[2], single[2], rtol=2e-3, atol=1e-5)
    d = np.abs(single[3] - multi[3])
    assert float(d.mean()) < 1e-4 and float(d.max()) <= 2 * 4 * 5e-4 + 1e-6, (float(d.mean()), float(d.max()))
