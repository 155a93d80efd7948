"""The training-loop body of the reference's ``nerf/train_nerf.py`` (lines 151-176) as ONE fused step.

    render_rays -> loss_coarse + loss_fine (MSE, optional 0.1 * alpha MSE) -> backward -> optimizer.step() -> LR decay

The reference runs this through autograd and ``torch.optim.Adam`` (about 400 kernel launches per step).  Here the same
arithmetic is an explicit sequence of libb2r kernels on flat fp32 buffers -- tensor-core forward with kept bf16
activations, composite forward / reverse, fused dgrad + wgrad, fused Adam with the learning-rate schedule on the device --
captured once in a CUDA graph and replayed.  Multi-GPU: every rank takes its shard of the batch and
ONE all-reduce sums the flat gradient bucket of both models between the two halves of the step.

The models stay ordinary ``nn.Module``s: their parameters are re-pointed to views of the flat buffer, so
``state_dict()`` / ``torch.save`` / ``render_image`` keep working on them unchanged (``train_nerf.py:177-199``).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from . import models, ops


class NerfTrainStep:
    """Fused training step for two NeRF (or two SirenNeRF) models (coarse + fine), bf16 tensor-core MLP arithmetic, fp32
    master weights.

    Arguments mirror the names train_nerf.py reads from its config: ``learning_rate``, ``learning_rate_decay`` (in
    thousands of steps, train_nerf.py:171), ``use_alpha``, the render settings and the batch size.  ``batch_size`` is the
    number of rays THIS rank renders per step; the loss is normalised by ``batch_size * world`` (the global batch)."""

    def __init__(self, coarse_model, fine_model, near, far, coarse_sample_num, fine_sample_num, batch_size, *,
                 learning_rate=5e-4, learning_rate_decay=0, use_alpha=False, betas=(0.9, 0.999), eps=1e-8, graph=True,
                 group=None):
        kind = models.model_kind(coarse_model)
        if kind not in (models.KIND_NERF, models.KIND_SIREN) or models.model_kind(fine_model) != kind:
            raise TypeError("NerfTrainStep needs two NeRF or two SirenNeRF models (nerf/nerf.py:52-94, 120-170; train_nerf.py:89-95)")
        self.kind = kind
        self.models = (coarse_model, fine_model)
        self.dev = next(coarse_model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("models must live on a CUDA device: the B200 training step has no CPU fallback")
        self.near, self.far = float(near), float(far)
        self.sc, self.sf, self.batch = int(coarse_sample_num), int(fine_sample_num), int(batch_size)
        self.lr0, self.decay_steps = float(learning_rate), float(learning_rate_decay) * 1000.0
        self.use_alpha, self.betas, self.eps = bool(use_alpha), betas, float(eps)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.use_graph = bool(graph)
        n = models.NERF_NUMEL if kind == models.KIND_NERF else models.SIREN_NUMEL
        self.n = n
        # flat fp32 master weights of both models; the nn.Parameters become views of it
        self.params = torch.empty((2 * n,), dtype=torch.float32, device=self.dev)
        for i, m in enumerate(self.models):
            off = i * n
            for p in models.param_list(m, kind):
                k = p.numel()
                self.params[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.params[off:off + k].view(p.shape)
                off += k
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.state = torch.zeros((4,), dtype=torch.float32, device=self.dev)
        # the two torch.linspace vectors are made on the host: their rounding is a contract
        self.z_lin = torch.linspace(self.near, self.far, steps=self.sc, device="cpu").to(self.dev)
        self.u = torch.linspace(0.0, 1.0, steps=self.sf, device="cpu").to(self.dev)
        b = self.batch
        self.in_rays = torch.zeros((b, 2, 3), dtype=torch.float32, device=self.dev)
        self.in_rgb = torch.zeros((b, 3), dtype=torch.float32, device=self.dev)
        self.in_alpha = torch.zeros((b,), dtype=torch.float32, device=self.dev)
        self.in_t = torch.zeros((b, self.sc), dtype=torch.float32, device=self.dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=self.dev)
        self.psnr = torch.zeros((), dtype=torch.float32, device=self.dev)
        self._graphs = None
        self._draw_t = True

    # ---- the step, as plain launches on the current stream -------------------------------------------------------
    def _forward_backward(self):
        n, kind = self.n, self.kind
        rays, rays_d = self.in_rays, self.in_rays[:, 1]
        bg = float(self.batch * self.world)
        if self._draw_t:
            self.in_t.copy_(torch.rand((self.batch, self.sc), device=self.dev))            # nerf/render.py:131
        flat_c, flat_f = self.params[:n], self.params[n:]
        z, mids = ops.stratified_z(self.z_lin, self.in_t)
        raw_c, saved_c = ops.tc_train_forward(ops.pack_tc(flat_c, kind), kind, rays, z)
        rgb_c, _, acc_c, w_c, ctx_c = ops.composite_forward(raw_c, z, rays_d, True)
        z_f = ops.sample_pdf(mids, w_c[:, 1:-1], self.sf, u=self.u, z_coarse=z, want_samples=False)["sorted"]
        raw_f, saved_f = ops.tc_train_forward(ops.pack_tc(flat_f, kind), kind, rays, z_f)
        rgb_f, _, acc_f, _, ctx_f = ops.composite_forward(raw_f, z_f, rays_d, False)
        # train_nerf.py:157-166
        e_c, e_f = rgb_c - self.in_rgb, rgb_f - self.in_rgb
        loss_c, loss_f = (e_c * e_c).sum() / (bg * 3), (e_f * e_f).sum() / (bg * 3)
        self.psnr.copy_(-10.0 * torch.log10((e_f * e_f).mean()))
        g_acc_c = g_acc_f = None
        if self.use_alpha:
            a_c, a_f = acc_c - self.in_alpha, acc_f - self.in_alpha
            loss_c = loss_c + 0.1 * (a_c * a_c).sum() / bg
            loss_f = loss_f + 0.1 * (a_f * a_f).sum() / bg
            g_acc_c, g_acc_f = a_c * (0.2 / bg), a_f * (0.2 / bg)
        self.loss.copy_(loss_f + loss_c)
        self.grads.zero_()
        d_raw_f = ops.composite_backward(ctx_f, e_f * (2.0 / (bg * 3)), None, g_acc_f)
        ops.tc_train_backward(ops.pack_tc_bwd(flat_f, kind), kind, raw_f, d_raw_f, saved_f, self.grads[n:])
        d_raw_c = ops.composite_backward(ctx_c, e_c * (2.0 / (bg * 3)), None, g_acc_c)
        ops.tc_train_backward(ops.pack_tc_bwd(flat_c, kind), kind, raw_c, d_raw_c, saved_c, self.grads[:n])

    def _optimize(self):
        ops.adam_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.state, self.lr0, 0.1, self.decay_steps,
                      self.betas, self.eps)

    def _capture(self):
        """warm up on a side stream, capture forward+backward and the optimiser as two graphs, restore the state the
        warm-up steps changed (weights, moments, step count; the RNG offset is not restored)."""
        keep = [t.clone() for t in (self.params, self.exp_avg, self.exp_avg_sq, self.state)]
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._forward_backward()
                self._optimize()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            self._forward_backward()
        with torch.cuda.graph(g2, pool=g1.pool()):
            self._optimize()
        for t, k in zip((self.params, self.exp_avg, self.exp_avg_sq, self.state), keep):
            t.copy_(k)
        self._graphs = (g1, g2)

    def __call__(self, batch_rays, batch_rgb, batch_alpha=None, t_rand=None):
        """One optimisation step on this rank's rays [batch,2,3] / colours [batch,3] (/ alpha [batch]).  Returns
        (loss, psnr) as 0-d CUDA tensors (global-batch loss share of this rank, fine-pass PSNR of this rank's rays);
        no host synchronisation happens here."""
        self.in_rays.copy_(torch.as_tensor(batch_rays).reshape(self.batch, 2, 3), non_blocking=True)
        self.in_rgb.copy_(torch.as_tensor(batch_rgb).reshape(self.batch, 3), non_blocking=True)
        if self.use_alpha:
            if batch_alpha is None:
                raise ValueError("use_alpha=True needs batch_alpha")
            self.in_alpha.copy_(torch.as_tensor(batch_alpha).reshape(self.batch), non_blocking=True)
        draw = t_rand is None
        if not draw:
            self.in_t.copy_(torch.as_tensor(t_rand).reshape(self.batch, self.sc), non_blocking=True)
        if self.use_graph:
            if self._graphs is None or draw != self._draw_t:
                self._draw_t = draw
                self._capture()
            self._graphs[0].replay()
        else:
            self._draw_t = draw
            self._forward_backward()
        if self.world > 1:
            dist.all_reduce(self.grads, group=self.group)                     # one 4.75 MB bucket, summed (loss is / global batch)
        if self.use_graph:
            self._graphs[1].replay()
        else:
            self._optimize()
        for m in self.models:                       # the weights changed behind torch's version counters
            ops.invalidate_packed(m)
        return self.loss, self.psnr

    # ---- bookkeeping ----------------------------------------------------------------------------------------------
    @property
    def global_step(self) -> int:
        return int(self.state[:1].view(torch.int32).item())

    @property
    def learning_rate(self) -> float:
        """learning rate the NEXT step will use (train_nerf.py:170-175)."""
        t = self.global_step
        return self.lr0 * (0.1 ** (t / self.decay_steps)) if self.decay_steps > 0 else self.lr0

    def state_dict(self) -> dict:
        return {"global_step": self.global_step, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "learning_rate": self.lr0, "learning_rate_decay": self.decay_steps / 1000.0}

    def load_state_dict(self, sd: dict) -> None:
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        t = int(sd["global_step"])
        b1, b2 = self.betas
        self.state.zero_()
        self.state[:1].view(torch.int32).fill_(t)
        if t > 0:
            self.state[1] = self.lr0 * (0.1 ** ((t - 1) / self.decay_steps)) if self.decay_steps > 0 else self.lr0
            self.state[2] = 1.0 - b1 ** t
            self.state[3] = math.sqrt(1.0 - b2 ** t)


class RayBatcher:
    """The GPU-resident training-ray buffer and the two batch samplers of ``nerf/train_nerf.py``:

    * lines 78-84: rays of every training pose + the pixels' rgba as ONE shuffled ``[N*H*W, 10]`` device tensor.  The
      reference builds it on the host with numpy ``get_rays`` (N x H x W x 6 floats) and uploads it; here the rays are
      generated on the device (``ops.raygen``, bit-exact origins, directions <= 1 ulp) and only the images cross PCIe.  The
      permutation is drawn with ``np.random.shuffle`` on an index vector, which consumes numpy's stream exactly like the
      reference's in-place shuffle of the rows and yields the same row order;
    * lines 125-137: the start-up sampler -- a random training image, the rays of its CENTRE crop (``get_rays`` of a
      half-size image with the full-size focal), ``batch_size`` pixels without replacement (``np.random.choice`` twice, in
      the reference's order);
    * lines 139-145: consecutive ``batch_size`` slices; at the end of an epoch the reference draws ``torch.randperm`` but
      assigns the shuffled copy to a misspelt name (``rays_rgb``), so the order never changes -- ``reshuffle=False`` (default)
      reproduces that (the draw still happens), ``reshuffle=True`` applies the permutation.

    ``next_batch`` / ``startup_batch`` return ``(rays[B,2,3], rgb[B,3], alpha[B])`` views for ``NerfTrainStep``; with a process
    group every rank draws the same global batch and keeps its ``rank``-th slice."""

    def __init__(self, images, poses, focal, batch_size, *, device=None, reshuffle=False, rank=0, world=1):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("RayBatcher keeps the ray buffer on a CUDA device: there is no CPU fallback")
        self.dev, self.focal, self.batch = dev, focal, int(batch_size)
        self.images = torch.as_tensor(images, dtype=torch.float32).to(dev)                     # [N,H,W,4]
        self.poses = [p for p in (poses.detach().cpu().numpy() if isinstance(poses, torch.Tensor) else poses)]
        n, h, w, _ = self.images.shape
        self.n, self.h, self.w = n, h, w
        rays = torch.cat([ops.raygen(w, h, focal, p[:3, :4], device=dev).reshape(-1, 6) for p in self.poses])      # [N*H*W,6]
        rows = torch.cat([rays, self.images.reshape(-1, 4)], 1)
        import numpy as np
        perm = np.arange(rows.shape[0])
        np.random.shuffle(perm)                                                                # same draws as shuffling the rows
        self.rays_rgba = rows[torch.from_numpy(perm).to(dev)].contiguous()
        self.batch_num = -(-rows.shape[0] // self.batch)
        self.batch_idx = 0
        self.reshuffle = bool(reshuffle)
        self.rank, self.world = int(rank), int(world)

    def _split(self, batch):
        if self.world > 1:
            per = batch.shape[0] // self.world
            batch = batch[self.rank * per:(self.rank + 1) * per]
        return batch[:, :6].reshape(-1, 2, 3), batch[:, 6:9], batch[:, 9]

    def startup_batch(self):
        import numpy as np
        sw, sh, left, top = int(self.w / 2), int(self.h / 2), int(self.w / 4), int(self.h / 4)
        i = np.random.choice(range(self.n))
        rays = ops.raygen(sw, sh, self.focal, self.poses[i][:3, :4], device=self.dev).reshape(-1, 6)
        rgba = self.images[i, top:top + sh, left:left + sw].reshape(-1, 4)
        idx = np.random.choice(range(sw * sh), size=self.batch, replace=False)
        return self._split(torch.cat([rays, rgba], 1)[torch.from_numpy(idx).to(self.dev)])

    def next_batch(self):
        b = self.rays_rgba[self.batch_idx * self.batch:(self.batch_idx + 1) * self.batch]
        self.batch_idx += 1
        if self.batch_idx == self.batch_num:
            shuffle_idx = torch.randperm(self.rays_rgba.shape[0])                              # train_nerf.py:143 (CPU generator)
            if self.reshuffle:
                self.rays_rgba = self.rays_rgba[shuffle_idx.to(self.dev)]
            self.batch_idx = 0
        return self._split(b)
