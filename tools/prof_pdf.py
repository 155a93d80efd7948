"""sample_pdf + merge at the headline shape (640,000 rays, 63 bins, 128 fine + 64 coarse samples) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msra_practice_project_b200 import ops
n, sc, sf = 640000, 64, 128
g = torch.Generator().manual_seed(0)
w = (torch.rand(n, sc, generator=g) ** 8).cuda()
# the inputs render_rays passes (nerf/render.py:126-142): stratified jittered coarse samples, bins = the strata mid-points
z_lin = torch.linspace(2.0, 6.0, sc)
mids_c = 0.5 * (z_lin[1:] + z_lin[:-1])
upper, lower = torch.cat([mids_c, z_lin[-1:]]), torch.cat([z_lin[:1], mids_c])
z = (lower + (upper - lower) * torch.rand(n, sc, generator=g)).cuda()
mids = mids_c.cuda()
u = torch.linspace(0.0, 1.0, sf, device="cpu").cuda()
for _ in range(3):
    out = ops.sample_pdf(mids, w[:, 1:-1], sf, u=u, z_coarse=z, want_samples=False)["sorted"]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = ops.sample_pdf(mids, w[:, 1:-1], sf, u=u, z_coarse=z, want_samples=False)["sorted"]
e1.record(); torch.cuda.synchronize()
print("sample_pdf + merge: %.3f ms" % (e0.elapsed_time(e1) / 10))
