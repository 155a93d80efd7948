"""GPU tests of the CUDA-graphed pi-GAN generator step (train_step.GeneratorStep; pi_GAN/train.py:121-145) and the two entry points
it added: b2r_raygen_poses (poses in device memory) and b2r_adam_step_floor (learning-rate floor)."""
import numpy as np
import pytest
import torch

from msra_practice_project_b200 import models, ops, pigan_render
from msra_practice_project_b200.train_step import GeneratorStep

pytestmark = pytest.mark.gpu


def test_raygen_poses_bit_identical_to_per_pose_raygen():
    """b2r_raygen_poses == b2r_raygen pose by pose, in the float32 mode (python-float focal) and the pi-GAN mode (np.float64 focal,
    pi_GAN/modules.py:127)."""
    poses = np.stack([pigan_render.camera_pos_to_transform_matrix(1.0, 0.3 * np.sin(i), 0.15 * np.cos(i)) for i in range(5)])
    assert poses.dtype == np.float32
    for w, h, focal in ((16, 12, 16 * 1.3875), (32, 32, np.float64(32 / 2 / np.tan(6 * np.pi / 180)))):
        ref = torch.cat([ops.raygen(w, h, focal, p) for p in poses])
        got = ops.raygen_poses(w, h, focal, torch.from_numpy(poses).cuda())
        assert torch.equal(ref, got)
        got3 = ops.raygen_poses(w, h, focal, torch.from_numpy(poses[:, :3]).cuda())
        assert torch.equal(ref, got3)


def test_adam_step_floor_schedule_matches_torch_adam():
    """b2r_adam_step_floor against torch.optim.Adam(betas=(0, 0.9)) driven with the schedule of pi_GAN/train.py:140-145."""
    torch.manual_seed(0)
    n, lr0, lr_end, decay = 1003, 5e-5, 1e-5, 3.0
    p0 = torch.randn(n, device="cuda")
    pad = (n + 3) // 4 * 4
    p = torch.zeros(pad, device="cuda"); p[:n] = p0
    g, m, v, st = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p), torch.zeros(4, device="cuda")
    q = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([q], lr=lr0, betas=(0.0, 0.9))
    for t in range(1, 7):
        grad = torch.randn(n, device="cuda") * (1.0 + t)
        g[:n] = grad
        ops.adam_step(p, g, m, v, st, lr0, 0.1, decay, (0.0, 0.9), 1e-8, lr_end=lr_end)
        q.grad = grad.clone()
        opt.step()
        for group in opt.param_groups:                                     # train.py:142-145, after the step
            group["lr"] = lr_end + (lr0 - lr_end) * (0.1 ** (t / decay))
        assert abs(float(st[1]) - (lr_end + (lr0 - lr_end) * 0.1 ** ((t - 1) / decay))) < 1e-11
    assert float((p[:n] - q.detach()).abs().max()) < 2e-7
    assert float(p[n:].abs().max()) == 0.0


def _make(seed=0, res=16, s=8):
    torch.manual_seed(seed)
    return models.Generator(256, res, near=0.5, far=1.5, fov=12, coarse_samples=s, fine_samples=s).cuda()


@pytest.mark.parametrize("graph", [False, True])
def test_generator_step_matches_autograd_and_torch_adam(graph):
    """Three generator updates through GeneratorStep (eager launches / CUDA-graph replay) against generator(z) + loss.backward() +
    torch.optim.Adam(betas=(0, 0.9)) with the train.py:140-145 schedule, on the same z / poses / jitter: same images every step
    (the weights feed back into the next step's render) and the same parameters at the end."""
    b, res, s = 2, 16, 8
    rs = np.random.RandomState(1)
    zs = [torch.from_numpy(rs.randn(b, 256).astype(np.float32)).cuda() for _ in range(3)]
    poses = [np.stack([pigan_render.camera_pos_to_transform_matrix(1.0, 0.3 * rs.randn(), 0.15 * rs.randn()) for _ in range(b)]) for _ in range(3)]
    ts = [torch.from_numpy(rs.rand(b, res * res, s).astype(np.float32)).cuda() for _ in range(3)]
    target = torch.from_numpy(rs.rand(b, 3, res, res).astype(np.float32)).cuda()
    lr0, lr_end, decay = 5e-4, 1e-4, 0.002          # lr_decay in thousands of steps: 2 steps per decade, so the schedule matters
    old = ops.set_grad_precision("bf16")
    try:
        ref = _make()
        opt = torch.optim.Adam(ref.parameters(), lr=lr0, betas=(0.0, 0.9))
        ref_imgs = []
        for t in range(3):
            img = ref(zs[t], poses=list(poses[t]), t_rand=ts[t])
            loss = ((img - target) ** 2).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
            for group in opt.param_groups:
                group["lr"] = lr_end + (lr0 - lr_end) * (0.1 ** ((t + 1) / (decay * 1000)))
            ref_imgs.append(img.detach().clone())
        gen = _make()
        p_start = [p.detach().clone() for p in gen.parameters()]
        step = GeneratorStep(gen, b, learning_rate=lr0, learning_rate_end=lr_end, lr_decay=decay, graph=graph)
        for t in range(3):
            img = step.forward(zs[t], poses[t], t_rand=ts[t])
            assert tuple(img.shape) == (b, 3, res, res) and not img.requires_grad
            diff = (img - ref_imgs[t]).abs()
            if t == 0:                               # same kernels on the same weights
                assert float(diff.max()) < 1e-5, float(diff.max())
            # later steps: Adam with beta1 = 0 moves every weight by ~lr * sign(g), so gradient entries at the atomics' noise level may
            # step the other way, and a ray whose LAST sample sits at sigma ~ 0 flips as a whole (nerf/render.py:92): bound the mean and
            # the share of pixels that moved, not the maximum
            assert float(diff.mean()) < 5e-3 and float((diff > 2e-2).float().mean()) < 0.02, (t, float(diff.mean()), float((diff > 2e-2).float().mean()))
            step.backward(2.0 * (img - target) / img.numel())
        assert step.global_step == 3
    finally:
        ops.set_grad_precision(old)
    for p in gen.parameters():                               # still views of the flat bucket
        assert step.params.data_ptr() <= p.data_ptr() < step.params.data_ptr() + step.params.numel() * 4
    da = torch.cat([(p.detach() - p0).double().reshape(-1) for p, p0 in zip(gen.parameters(), p_start)])
    db = torch.cat([(q.detach() - p0).double().reshape(-1) for q, p0 in zip(ref.parameters(), p_start)])
    cos = float((da * db).sum() / (da.norm() * db.norm()))
    assert float(da.norm()) > 0 and cos > 0.98, cos
    print("GeneratorStep graph=%s: update cosine vs autograd + torch Adam %.5f, |update| %.3e vs %.3e" % (graph, cos, float(da.norm()), float(db.norm())))


def test_generator_step_drawn_jitter_and_poses_change_between_replays():
    """Default mode (jitter drawn by the captured torch.rand, poses drawn on the host and copied into device memory): two replays with the
    same z give different images (new jitter / poses reach the graph), finite, in [0, 1]; state_dict keys are the reference's."""
    gen = _make(seed=1)
    keys = set(gen.state_dict().keys())
    step = GeneratorStep(gen, 2)
    torch.autograd.set_detect_anomaly(True)                 # what importing the reference's nerf/nerf.py leaves behind (line 2)
    z = torch.randn(2, 256, device="cuda")
    np.random.seed(0)
    a = step.forward(z).clone()
    step.backward(torch.full_like(a, 1e-3))
    b = step.forward(z).clone()
    assert torch.isfinite(a).all() and torch.isfinite(b).all() and float(a.min()) >= 0.0 and float(a.max()) <= 1.0
    assert float((a - b).abs().max()) > 1e-4
    pose_fixed = np.stack([pigan_render.camera_pos_to_transform_matrix(1.0, 0.1, 0.05)] * 2)
    c = step.forward(z, pose_fixed).clone()
    d = step.forward(z, pose_fixed).clone()
    assert float((c - d).abs().max()) > 1e-6                 # same z, same poses, no update in between: only the jitter differs
    torch.autograd.set_detect_anomaly(False)
    assert set(gen.state_dict().keys()) == keys and step.global_step == 1


def test_generator_step_recaptures_on_resolution_change():
    """Progressive growing (pi_GAN/train.py:150-160 calls generator.set_resolution between stages): the captured graphs are keyed by
    the renderer's settings, so a new resolution re-captures and keeps training the same flat parameter buffer."""
    gen = _make(seed=2, res=16, s=8)
    step = GeneratorStep(gen, 2, learning_rate=1e-4)
    z = torch.randn(2, 256, device="cuda")
    a = step.forward(z)
    assert tuple(a.shape) == (2, 3, 16, 16)
    step.backward(torch.full_like(a, 1e-3))
    before = step.params.clone()
    gen.set_resolution(32)                                   # focal follows the width (modules.py:141)
    b = step.forward(z)
    assert tuple(b.shape) == (2, 3, 32, 32) and torch.isfinite(b).all()
    step.backward(torch.full_like(b, 1e-3))
    assert step.global_step == 2 and float((step.params - before).abs().max()) > 0
    with pytest.raises(ValueError):
        step.backward(torch.zeros(2, 3, 16, 16, device="cuda"))
    with pytest.raises(ValueError):
        step.forward(torch.randn(3, 256, device="cuda"))
