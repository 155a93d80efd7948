"""cProfile of pigan_render.render_batch (64 latents x 128x128, 24+24): where the host time of the step goes."""
import os, sys, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from msra_practice_project_b200 import models, pigan_render
n_lat, res, s_ = 64, 128, 24
torch.manual_seed(0)
net = models.FilmSirenNeRF().cuda()
g = torch.Generator().manual_seed(0)
film = torch.cat([1.0 + 0.2 * torch.randn(n_lat, 9, 256, generator=g), 0.1 * torch.randn(n_lat, 9, 256, generator=g)], -1).cuda()
focal = np.float64(res / 2 / np.tan(6 * np.pi / 180))
poses = [pigan_render.camera_pos_to_transform_matrix(1, 0.3 * np.sin(i), 0.15 * np.cos(i)) for i in range(n_lat)]
def step():
    with torch.no_grad():
        return pigan_render.render_batch(net, film, poses, res, res, focal, 0.5, 1.5, s_, s_)
for _ in range(3): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host launch time %.2f ms, until GPU done %.2f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
pr = cProfile.Profile(); pr.enable(); step(); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
