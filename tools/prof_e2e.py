"""Where the end-to-end frame time goes beyond the device time (render_image: host pose in, numpy images out)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from msra_practice_project_b200 import models, nerf_render, pigan_render
dev = torch.device("cuda", 0)
W = H = 800; sc, sf = 64, 128
torch.manual_seed(0)
coarse, fine = models.NeRF().to(dev), models.NeRF().to(dev)
pose = pigan_render.camera_pos_to_transform_matrix(4.0, 0.3, -30 * np.pi / 180)
focal = W * 1.3875
def t(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = t()
    with torch.no_grad():
        packed = torch.empty((H * W, 5), device=dev)
        tr = nerf_render._draw_t_rand(W * H, sc, 1024 * 16, dev)
        t1 = t()
        nerf_render.render_image_device(W, H, focal, pose, 2.0, 6.0, coarse, fine, sc, sf, t_rand=tr, fine_out=packed)
        t2 = t()
        out = nerf_render.maps_to_numpy(packed, H, W)
        t3 = t()
    print("iter %d: jitter %.2f ms, render %.2f ms, read-back %.2f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
# read-back pieces
n = H * W
stage = torch.empty((n * 5,), dtype=torch.float32, device="cpu", pin_memory=True)
for it in range(3):
    a0 = t(); flat = torch.cat([packed[:, :3].reshape(-1), packed[:, 3], packed[:, 4]]); a1 = t()
    stage.copy_(flat, non_blocking=True); a2 = t()
    a = stage.numpy(); r = (a[:3 * n].copy(), a[3 * n:4 * n].copy(), a[4 * n:].copy()); a3 = time.perf_counter()
    print("  de-interleave %.2f ms, D2H %.2f ms, 3 host copies %.2f ms" % ((a1 - a0) * 1e3, (a2 - a1) * 1e3, (a3 - a2) * 1e3))
for it in range(3):
    t0 = t(); nerf_render.render_image(W, H, focal, pose, 2.0, 6.0, coarse, fine, sc, sf); t1 = t()
    print("render_image: %.2f ms" % ((t1 - t0) * 1e3))
print("host cores", os.cpu_count())
