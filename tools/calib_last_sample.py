"""Calibration of the last-sample sign check (ops._LAST_REL): for each model kind, the number of rays whose sign(sigma_last)
still differs from the fp32 path, and the fraction of rays sent to the fp32 engine, as a function of the band `rel`.
Run on a B200:  python tools/calib_last_sample.py > gpurun_out/calib_last_sample.txt"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msra_practice_project_b200 import models, ops, pigan_render  # noqa: E402


def sweep(name, kind, model, rays, z, run):
    n = z.shape[0]
    with torch.no_grad():
        lb = ops.mlp(model, rays=rays, z=z[:, -1:].contiguous(), precision="fp32")[:, 3] if run is None else run("fp32")
    print(f"== {name}: {n} rays, S = {z.shape[1]}; fp32 sigma_last > 0 on {100.0 * float((lb > 0).float().mean()):.1f} % of rays")
    old = ops._LAST_REL[kind]
    for e in (-14, -12, -11, -10, -9, -8, -7, -6, -5):
        ops.set_last_sample_band(kind, 2.0 ** e)
        before = ops.last_sample_stats["flagged"]
        with torch.no_grad():
            a = (ops.mlp(model, rays=rays, z=z, precision="bf16") if run is None else run("bf16")).view(n, -1, 4)[:, -1, 3]
        flagged = ops.last_sample_stats["flagged"] - before
        flips = int(((a > 0) != (lb > 0)).sum())
        print(f"   rel = 2^{e:<4d} flagged {flagged:8d} ({100.0 * flagged / n:6.3f} %)   sign flips left {flips}")
    ops.set_last_sample_band(kind, old)


def main():
    torch.cuda.set_device(0)
    pose = pigan_render.camera_pos_to_transform_matrix(4.0, 0.3, -30 * np.pi / 180)
    rays = ops.raygen(800, 800, 800 * 1.3875, pose)
    torch.manual_seed(5)
    z, _ = ops.stratified_z(torch.linspace(2.0, 6.0, 64).cuda(), torch.rand(rays.shape[0], 64, device="cuda"))
    torch.manual_seed(0)
    c, f = models.NeRF().cuda(), models.NeRF().cuda()
    sweep("NeRF Xavier (coarse model, bench weights)", models.KIND_NERF, c, rays, z, None)
    sweep("NeRF Xavier (fine model)", models.KIND_NERF, f, rays, z, None)
    torch.manual_seed(0)
    d = models.damp_nerf_(models.NeRF()).cuda()
    sweep("NeRF damped field", models.KIND_NERF, d, rays, z, None)
    with torch.no_grad():                                       # a "trained-like" scale: 4x larger trunk activations / sigma weights
        big = models.NeRF().cuda()
        for p in big.parameters():
            p.mul_(1.3)
    sweep("NeRF Xavier x 1.3 per layer", models.KIND_NERF, big, rays, z, None)
    torch.manual_seed(0)
    sn = models.SirenNeRF().cuda()
    sweep("SirenNeRF default init", models.KIND_SIREN, sn, rays[:320000], z[:320000].contiguous(), None)
    # FiLM-SIREN, 16 latents x 128 x 128, batched
    torch.manual_seed(0)
    m = models.FilmSirenNeRF().cuda()
    g = torch.Generator().manual_seed(0)
    nl, res = 16, 128
    films = torch.cat([1.0 + 0.2 * torch.randn(nl, 9, 256, generator=g), 0.1 * torch.randn(nl, 9, 256, generator=g)], -1).cuda()
    focal = np.float64(res / 2 / np.tan(6 * np.pi / 180))
    rays_f = torch.cat([ops.raygen(res, res, focal, pigan_render.camera_pos_to_transform_matrix(1, 0.3 * np.sin(i), 0.15 * np.cos(i))) for i in range(nl)])
    zf, _ = ops.stratified_z(torch.linspace(0.5, 1.5, 24).cuda(), torch.rand(rays_f.shape[0], 24, device="cuda"))
    n1 = res * res

    def run(prec):
        if prec == "bf16":
            return ops.mlp_film_batched(m, films, rays_f, zf, n1 * 24)
        outs = []
        for b in range(nl):
            m.set_film_params(films[b])
            outs.append(ops.mlp(m, rays=rays_f[b * n1:(b + 1) * n1], z=zf[b * n1:(b + 1) * n1, -1:].contiguous(), precision="fp32")[:, 3])
        return torch.cat(outs)
    sweep("FiLM-SIREN, 16 latents x 128x128 (batched)", models.KIND_FILM, m, rays_f, zf, run)


if __name__ == "__main__":
    main()
