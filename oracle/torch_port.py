"""CPU port of the reference render path on torch ATen ops (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference's arithmetic library for this path is PyTorch ATen: on a CPU it runs
MKL sgemm + vectorised elementwise kernels on all host threads.  ``render_oracle.py`` (numpy) is the
bit-level checker; this file restates the same algorithm with the same ATen calls the reference makes,
so that a CPU baseline can be timed at the speed the reference itself would reach on the box's host
cores.  ``bench.py`` (``cpu_baseline`` and ``--impl reference``) times the UNMODIFIED reference from the
git-ignored copy ``oracle/_ref`` (tools/install_ref.sh, oracle/ref_loader.py) and falls back to this
port (``kind: "port"``) only when that copy is absent; the port is also the "stock eager PyTorch on the
same B200" incumbent.

Only ``tests/``, ``__graft_entry__.smoke()`` and bench.py's CPU arm may import this module.  It is
pinned by tests/test_oracle_golden.py against the golden outputs of the unmodified reference.
Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import torch


def posenc(x: torch.Tensor, length: int) -> torch.Tensor:
    """nerf/nerf.py:44-49: [sin(2^i x), cos(2^i x)] blocks, i = 0..L-1, no identity term."""
    out = []
    for i in range(length):
        out.append(torch.sin(x * 2.0 ** i))
        out.append(torch.cos(x * 2.0 ** i))
    return torch.cat(out, -1)


def nerf_mlp(sd: dict, x: torch.Tensor) -> torch.Tensor:
    """nerf/nerf.py:75-94 with the state-dict tensors ``sd`` (reference key names)."""
    lin = torch.nn.functional.linear
    pe, de = posenc(x[:, :3], 10), posenc(x[:, 3:], 4)
    h = pe
    for l in range(8):
        if l == 5:
            h = torch.cat([pe, h], -1)                                      # nerf.py:84
        h = torch.relu(lin(h, sd[f"layers_pos.{l}.weight"], sd[f"layers_pos.{l}.bias"]))
    sigma = torch.relu(lin(h, sd["output_layer_sigma.weight"], sd["output_layer_sigma.bias"]))
    h = lin(h, sd["layers_dir.0.weight"], sd["layers_dir.0.bias"])
    h = torch.relu(lin(torch.cat([h, de], -1), sd["layers_dir.1.weight"], sd["layers_dir.1.bias"]))
    rgb = torch.sigmoid(lin(h, sd["output_layer_rgb.weight"], sd["output_layer_rgb.bias"]))
    return torch.cat([rgb, sigma], -1)


def run_network(pts: torch.Tensor, view_dirs: torch.Tensor, net, chunk: int = 1024 * 64) -> torch.Tensor:
    """nerf/render.py:59-75: flatten, broadcast view dirs, evaluate in 65,536-row chunks."""
    n, s, _ = pts.shape
    x = torch.cat([pts.reshape(-1, 3), view_dirs[:, None].expand(n, s, 3).reshape(-1, 3)], -1)
    return torch.cat([net(x[i:i + chunk]) for i in range(0, x.shape[0], chunk)], 0).reshape(n, s, 4)


def raw_to_outputs(raw: torch.Tensor, z: torch.Tensor, rays_d: torch.Tensor):
    """nerf/render.py:78-103."""
    dists = z[:, 1:] - z[:, :-1]
    dists = torch.cat([dists, torch.full((z.shape[0], 1), 1e10)], -1) * torch.norm(rays_d, dim=-1, keepdim=True)
    alpha = 1.0 - torch.exp(-raw[..., 3] * dists)
    trans = torch.cumprod(torch.cat([torch.ones((z.shape[0], 1)), 1.0 - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * trans
    acc = torch.sum(w, -1)
    rgb = torch.sum(w[..., None] * raw[..., :3], -2) + (1.0 - acc[..., None])
    return rgb, torch.sum(w * z, -1), acc, w


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_samples: int) -> torch.Tensor:
    """nerf/render.py:27-56."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)
    u = torch.linspace(0.0, 1.0, steps=n_samples).expand(cdf.shape[0], n_samples).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bins = bins.expand(cdf.shape[0], cdf.shape[-1])
    bins_b, bins_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return bins_b + (u - cdf_b) / denom * (bins_a - bins_b)


def render_rays(rays: torch.Tensor, near: float, far: float, coarse_net, fine_net, sc: int, sf: int, t_rand: torch.Tensor):
    """nerf/render.py:106-147 with the jitter passed in (the reference draws it with torch.rand, :131)."""
    rays_o, rays_d = rays[:, 0], rays[:, 1]
    view_dirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    z = torch.linspace(near, far, steps=sc).expand(rays.shape[0], sc)
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    upper = torch.cat([mids, z[:, -1:]], -1)
    lower = torch.cat([z[:, :1], mids], -1)
    z = lower + (upper - lower) * t_rand
    raw = run_network(rays_o[:, None] + rays_d[:, None] * z[..., None], view_dirs, coarse_net)
    rgb_c, depth_c, acc_c, w = raw_to_outputs(raw, z, rays_d)
    z_samples = sample_pdf(mids, w[:, 1:-1], sf)
    z, _ = torch.sort(torch.cat([z, z_samples], -1), -1)
    raw = run_network(rays_o[:, None] + rays_d[:, None] * z[..., None], view_dirs, fine_net)
    rgb_f, depth_f, acc_f, _ = raw_to_outputs(raw, z, rays_d)
    return rgb_c, depth_c, acc_c, rgb_f, depth_f, acc_f
