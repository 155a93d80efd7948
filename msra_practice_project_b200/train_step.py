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
        if self.batch < 1:
            raise ValueError("batch_size must be >= 1 ray per rank")
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
        # a short last batch (train_nerf.py:139-150 trains on whatever the final slice of an epoch holds) runs through the same fixed-shape
        # launch sequence: the missing rays are padding with loss weight 0, and the loss is normalised by the true global ray count, which
        # lives on the device (counts = [1 / global rays, 1 / this rank's rays]) so that the captured graphs replay unchanged
        self.ray_w = torch.ones((b,), dtype=torch.float32, device=self.dev)
        self.counts = torch.tensor([1.0 / (b * self.world), 1.0 / b], dtype=torch.float32, device=self.dev)
        self._n_valid, self._global = b, b * self.world
        self.sums = torch.zeros((4,), dtype=torch.float32, device=self.dev)      # [fine rgb, fine acc, coarse rgb, coarse acc] squared errors
        self.loss = torch.zeros((), dtype=torch.float32, device=self.dev)
        self.psnr = torch.zeros((), dtype=torch.float32, device=self.dev)
        self._graphs = None
        self._draw_t = True
        self._side = torch.cuda.Stream(device=self.dev) if self.world > 1 else None

    # ---- the step, as plain launches on the current stream -------------------------------------------------------
    # 17 launches of libb2r kernels + 2 fills + the jitter draw: stratified z, 2 x (pack, fused forward), composite (coarse: its weights feed
    # sample_pdf), sample_pdf + merge, 2 x (loss + composite reverse in one kernel, transposed pack, dgrad, wgrad, heads), loss finish, Adam.
    # The loss / PSNR / upstream-gradient arithmetic of train_nerf.py:157-166 (about 40 elementwise and reduction launches under autograd)
    # lives inside b2r_composite_loss_bwd.
    def _forward_and_fine_backward(self):
        n, kind = self.n, self.kind
        rays, rays_d = self.in_rays, self.in_rays[:, 1]
        if self._draw_t:
            self.in_t.copy_(torch.rand((self.batch, self.sc), device=self.dev))            # nerf/render.py:131
        flat_c, flat_f = self.params[:n], self.params[n:]
        z, mids = ops.stratified_z(self.z_lin, self.in_t)
        raw_c, saved_c = ops.tc_train_forward(ops.pack_tc(flat_c, kind), kind, rays, z)
        _, _, _, w_c, _ = ops.composite_forward(raw_c, z, rays_d, True)
        z_f = ops.sample_pdf(mids, w_c[:, 1:-1], self.sf, u=self.u, z_coarse=z, want_samples=False)["sorted"]
        raw_f, saved_f = ops.tc_train_forward(ops.pack_tc(flat_f, kind), kind, rays, z_f)
        self.grads.zero_()
        self.sums.zero_()
        alpha_t = self.in_alpha if self.use_alpha else None
        aw = 0.1 if self.use_alpha else 0.0                                               # train_nerf.py:161-163
        d_raw_f = ops.composite_loss_backward(raw_f, z_f, rays_d, self.in_rgb, alpha_t, self.ray_w, self.counts[:1], aw, self.sums[:2])
        ops.tc_train_backward(ops.pack_tc_bwd(flat_f, kind), kind, raw_f, d_raw_f, saved_f, self.grads[n:])
        self._coarse = (raw_c, z, saved_c, alpha_t, aw)

    def _coarse_backward(self):
        n, kind = self.n, self.kind
        raw_c, z, saved_c, alpha_t, aw = self._coarse
        d_raw_c = ops.composite_loss_backward(raw_c, z, self.in_rays[:, 1], self.in_rgb, alpha_t, self.ray_w, self.counts[:1], aw, self.sums[2:])
        ops.tc_train_backward(ops.pack_tc_bwd(self.params[:n], kind), kind, raw_c, d_raw_c, saved_c, self.grads[:n])
        ops.train_loss_finish(self.sums, self.counts[:1], aw, self.counts[1:], self.loss, self.psnr)

    def _forward_backward(self):
        self._forward_and_fine_backward()
        self._coarse_backward()

    def _optimize(self):
        ops.adam_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.state, self.lr0, 0.1, self.decay_steps,
                      self.betas, self.eps)

    def _capture(self):
        """warm up on a side stream, capture the step as three graphs -- forward + fine backward | coarse backward | optimiser: the seams
        are where the two gradient all-reduces go -- and restore the state the warm-up steps changed (weights, moments, step count; the
        RNG offset is not restored)."""
        keep = [t.clone() for t in (self.params, self.exp_avg, self.exp_avg_sq, self.state)]
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._forward_backward()
                self._optimize()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            self._forward_and_fine_backward()
        with torch.cuda.graph(g2, pool=g1.pool()):
            self._coarse_backward()
        with torch.cuda.graph(g3, pool=g1.pool()):
            self._optimize()
        for t, k in zip((self.params, self.exp_avg, self.exp_avg_sq, self.state), keep):
            t.copy_(k)
        self._graphs = (g1, g2, g3)

    def _set_valid(self, n_valid: int, global_count: int | None):
        g = n_valid * self.world if global_count is None else int(global_count)
        if (n_valid, g) == (self._n_valid, self._global):
            return
        self.ray_w[:n_valid] = 1.0
        self.ray_w[n_valid:] = 0.0
        self.counts.copy_(torch.tensor([1.0 / max(g, 1), 1.0 / max(n_valid, 1)], dtype=torch.float32), non_blocking=False)
        self._n_valid, self._global = n_valid, g

    def __call__(self, batch_rays, batch_rgb, batch_alpha=None, t_rand=None, global_count=None):
        """One optimisation step on this rank's rays [n,2,3] / colours [n,3] (/ alpha [n]), n <= batch_size (a short last batch is padded
        with zero-weight rays; ``global_count`` = the rays of the step over all ranks when the ranks' shares differ, RayBatcher.global_count).
        Returns (loss, psnr) as 0-d CUDA tensors (global-batch loss share of this rank, fine-pass PSNR of this rank's rays); no host
        synchronisation happens here."""
        batch_rays = torch.as_tensor(batch_rays)
        n_valid = int(batch_rays.reshape(-1, 2, 3).shape[0])
        if n_valid > self.batch or n_valid < 1:
            raise ValueError(f"this step was built for at most {self.batch} rays per rank, got {n_valid}")
        self._set_valid(n_valid, global_count)
        self.in_rays[:n_valid].copy_(batch_rays.reshape(n_valid, 2, 3), non_blocking=True)
        if n_valid < self.batch:
            # the padding must be a VALID ray (a zero direction would put NaNs into the MLP, and 0 * NaN into the gradients)
            self.in_rays[n_valid:] = self.in_rays[0]
        self.in_rgb[:n_valid].copy_(torch.as_tensor(batch_rgb).reshape(n_valid, 3), non_blocking=True)
        if self.use_alpha:
            if batch_alpha is None:
                raise ValueError("use_alpha=True needs batch_alpha")
            self.in_alpha[:n_valid].copy_(torch.as_tensor(batch_alpha).reshape(n_valid), non_blocking=True)
        draw = t_rand is None
        if not draw:
            self.in_t[:n_valid].copy_(torch.as_tensor(t_rand).reshape(n_valid, self.sc), non_blocking=True)
        if self.use_graph and (self._graphs is None or draw != self._draw_t):
            self._draw_t = draw
            self._capture()
        self._draw_t = draw
        main = torch.cuda.current_stream(self.dev)
        n = self.n
        # fine model first: its gradient half is all-reduced on a side stream while the coarse model's reverse mode runs
        if self.use_graph:
            self._graphs[0].replay()
        else:
            self._forward_and_fine_backward()
        if self.world > 1:
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                dist.all_reduce(self.grads[n:], group=self.group)                 # 2.4 MB, summed (the loss is / global count)
        if self.use_graph:
            self._graphs[1].replay()
        else:
            self._coarse_backward()
        if self.world > 1:
            dist.all_reduce(self.grads[:n], group=self.group)
            main.wait_stream(self._side)
        if self.use_graph:
            self._graphs[2].replay()
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
        if self.batch % int(world) != 0:
            raise ValueError(f"batch_size ({self.batch}) must be a multiple of the number of ranks ({world}): NerfTrainStep(batch_size=batch // world)")
        self.global_count = self.batch
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
        """this rank's share of a global batch: contiguous slices of ceil(m / world) rows (the last ranks' may be shorter or empty for
        the short final batch of an epoch); ``global_count`` = m, what the step's loss is normalised by."""
        self.global_count = int(batch.shape[0])
        if self.world > 1:
            per = -(-batch.shape[0] // self.world)
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


class GeneratorStep:
    """The generator update of ``pi_GAN/train.py:121-145`` as CUDA graphs around the caller's discriminator (which is out of scope
    here):

        images = step.forward(z)                       # graph 1: mapping network -> FiLM -> all latents rendered in one launch sequence
        g_loss = loss_f(discriminator(images)).mean()  # the caller's code, eager
        step.backward(torch.autograd.grad(g_loss, images)[0])    # graph 2: composite reverse -> fused FiLM-SIREN reverse mode ->
                                                       # d film -> mapping network;  [all-reduce];  graph 3: fused Adam + LR schedule

    Under autograd the same step is ~170 launches issued from Python (generator(z) + loss.backward() + torch.optim.Adam over 50
    parameter tensors): at the reference's early-stage shapes the GPU finishes before the host has queued the launches.  Here the
    launch sequence is captured once per (batch, resolution, sample counts) and replayed; everything that changes between steps
    lives in device buffers: z, the camera poses (b2r_raygen_poses reads them from device memory), the jitter (drawn by the
    captured torch.rand, whose Philox offset torch advances per replay), the upstream gradient, the step count and the decayed
    learning rate (b2r_adam_step_floor: lr_end + (lr0 - lr_end) * 0.1^((t-1) / (lr_decay * 1000)), train.py:140-145).

    The generator stays an ordinary ``nn.Module`` (``models.Generator`` or the reference's own class -- same attributes): its
    parameters become views of one flat fp32 buffer (state_dict() / torch.save keep working), Adam's moments are flat too.
    Multi-GPU: every rank runs its slice of the latent batch; ``backward`` all-reduces the flat gradient bucket once (SUM; scale the
    upstream gradient for the global batch, as g_loss.mean() over the global batch does)."""

    def __init__(self, generator, batch_size, *, learning_rate=5e-5, learning_rate_end=1e-5, lr_decay=500, betas=(0.0, 0.9), eps=1e-8,
                 graph=True, group=None):
        self.gen = generator
        net = generator.film_siren_nerf
        if models.model_kind(net) != models.KIND_FILM:
            raise TypeError("GeneratorStep needs a pi-GAN Generator (film_siren_nerf + mapping_network, pi_GAN/modules.py:164-197)")
        ps = [p for p in generator.parameters()]
        self.dev = ps[0].device
        if self.dev.type != "cuda":
            raise RuntimeError("the generator must live on a CUDA device: the B200 training step has no CPU fallback")
        self.batch = int(batch_size)
        if self.batch < 1:
            raise ValueError("batch_size must be >= 1 latent per rank")
        self.lr0, self.lr_end, self.decay_steps = float(learning_rate), float(learning_rate_end), float(lr_decay) * 1000.0
        self.betas, self.eps = betas, float(eps)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.use_graph = bool(graph)
        n = sum(p.numel() for p in ps)
        n_pad = (n + 3) // 4 * 4
        self.n = n
        self.params = torch.zeros((n_pad,), dtype=torch.float32, device=self.dev)
        off = 0
        for p in ps:                                     # generator.parameters() order = the order torch.optim.Adam would see
            k = p.numel()
            self.params[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.params[off:off + k].view(p.shape)
            off += k
        self._plist = ps
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.state = torch.zeros((4,), dtype=torch.float32, device=self.dev)
        self.z = torch.zeros((self.batch, int(generator.input_dim)), dtype=torch.float32, device=self.dev)
        self.poses = torch.zeros((self.batch, 4, 4), dtype=torch.float32, device=self.dev)
        self._poses_host = torch.zeros((self.batch, 4, 4), dtype=torch.float32, device="cpu", pin_memory=True)
        self.d_images = None
        self.images = None
        self.t_rand = None                               # injected jitter [B, H*W, coarse_samples] (parity runs); None: drawn per step
        self._draw_t = True
        self._graphs, self._key = None, None

    # ---- the three pieces, as plain launches on the current stream ---------------------------------------------------
    def _fwd(self):
        with torch.enable_grad():
            self.images = _generator_forward(self.gen, self.z, self.poses, None if self._draw_t else self.t_rand)
        if self.d_images is None or self.d_images.shape != self.images.shape:
            self.d_images = torch.zeros_like(self.images)

    def _bwd(self):
        gs = torch.autograd.grad([self.images], self._plist, [self.d_images], allow_unused=True)
        flat = [(g if g is not None else torch.zeros_like(p)).reshape(-1) for g, p in zip(gs, self._plist)]
        torch.cat(flat, out=self.grads[:self.n])

    def _opt(self):
        ops.adam_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.state, self.lr0, 0.1, self.decay_steps, self.betas,
                      self.eps, lr_end=self.lr_end)

    def _shape_key(self):
        r = self.gen.renderer
        return (int(r.width), int(r.height), int(r.coarse_samples), int(r.fine_samples), float(r.near), float(r.far), float(r.focal), self._draw_t)

    def _capture(self):
        """warm up on a side stream (this also fills the host-made linspace caches, so the capture sees no host-to-device copy), capture
        forward | backward | optimiser, and restore what the warm-up steps changed (weights, moments, step count)."""
        keep = [t.clone() for t in (self.params, self.exp_avg, self.exp_avg_sq, self.state)]
        # anomaly mode (the reference switches it on globally when nerf/nerf.py is imported, line 2) checks every backward result for NaNs
        # on the host, which invalidates a stream capture: the captured passes run with it off
        with torch.autograd.set_detect_anomaly(False):
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._fwd()
                    self.d_images.fill_(1e-3)
                    self._bwd()
                    self._opt()
            torch.cuda.current_stream(self.dev).wait_stream(s)
            g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            self.d_images = torch.zeros_like(self.images)     # allocated OUTSIDE the graphs' pool: the caller's gradient is copied into it
            with torch.cuda.graph(g1):
                self._fwd()
            with torch.cuda.graph(g2, pool=g1.pool()):
                self._bwd()
            with torch.cuda.graph(g3, pool=g1.pool()):
                self._opt()
        for t, k in zip((self.params, self.exp_avg, self.exp_avg_sq, self.state), keep):
            t.copy_(k)
        self._graphs = (g1, g2, g3)

    # ---- public API --------------------------------------------------------------------------------------------------
    def forward(self, z=None, poses=None, t_rand=None) -> torch.Tensor:
        """images [B,3,H,W] for this rank's latents z[B,input_dim] (None: drawn with torch.randn like train.py:126) and camera poses
        [B,4,4] (None: drawn with np.random in the reference's order, modules.py:153-158).  The returned tensor is a static buffer that the
        next forward() overwrites; it is detached -- differentiate the caller's loss with respect to it and hand the result to backward()."""
        import numpy as np
        if z is None:
            z = torch.randn(self.batch, self.z.shape[1], device=self.dev)
        if tuple(z.shape) != tuple(self.z.shape):
            raise ValueError(f"this step was built for z of shape {tuple(self.z.shape)}, got {tuple(z.shape)}")
        self.z.copy_(z, non_blocking=True)
        if poses is None:
            poses = np.stack([self.gen.renderer.draw_pose() for _ in range(self.batch)])
        if isinstance(poses, torch.Tensor) and poses.is_cuda:
            self.poses.copy_(poses.reshape(self.batch, 4, 4))
        else:
            self._poses_host.copy_(torch.as_tensor(np.asarray(poses, dtype=np.float32)).reshape(self.batch, 4, 4))
            self.poses.copy_(self._poses_host, non_blocking=True)
        self._draw_t = t_rand is None
        if t_rand is not None:                           # parity runs inject the jitter the reference draws with torch.rand (render.py:131)
            r = self.gen.renderer
            shape = (self.batch, int(r.width) * int(r.height), int(r.coarse_samples))
            if self.t_rand is None or tuple(self.t_rand.shape) != shape:
                self.t_rand = torch.zeros(shape, dtype=torch.float32, device=self.dev)
                self._graphs = None
            self.t_rand.copy_(torch.as_tensor(t_rand).reshape(shape), non_blocking=True)
        if self.use_graph:
            key = self._shape_key()
            if self._graphs is None or key != self._key:
                self._key = key
                self._capture()
            self._graphs[0].replay()
        else:
            self._fwd()
        return self.images.detach()

    def backward(self, d_images: torch.Tensor) -> None:
        """d loss / d images [B,3,H,W] of the images the last forward() returned -> gradients -> (all-reduce) -> Adam step."""
        if self.images is None:
            raise RuntimeError("backward() needs a forward() first")
        if tuple(d_images.shape) != tuple(self.images.shape):
            raise ValueError(f"d_images must have the images' shape {tuple(self.images.shape)}, got {tuple(d_images.shape)}")
        self.d_images.copy_(d_images, non_blocking=True)
        if self.use_graph:
            self._graphs[1].replay()
        else:
            self._bwd()
        if self.world > 1:
            dist.all_reduce(self.grads, group=self.group)
        if self.use_graph:
            self._graphs[2].replay()
        else:
            self._opt()

    @property
    def global_step(self) -> int:
        return int(self.state[:1].view(torch.int32).item())


def _generator_forward(gen, z, poses, t_rand):
    """Generator.forward (pi_GAN/modules.py:176-184) -- also for the reference's own class (same sub-modules): mapping network, then all
    latents through render_batch with the poses read from device memory."""
    from . import pigan_render
    film = gen.mapping_network(z)
    r = gen.renderer
    return pigan_render.render_batch(gen.film_siren_nerf, film, poses, r.width, r.height, r.focal, r.near, r.far, r.coarse_samples,
                                     r.fine_samples, t_rand=t_rand)
