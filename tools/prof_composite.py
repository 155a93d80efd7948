"""composite forward (coarse with weights, fine without) and stratified z at the headline shape (640,000 rays) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msra_practice_project_b200 import ops
n, sc, sf = 640000, 64, 128
g = torch.Generator().manual_seed(0)
d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda() * 1.1
z_lin = torch.linspace(2.0, 6.0, sc).cuda()
t = torch.rand(n, sc, generator=g).cuda()
raw_c = torch.rand(n, sc, 4, generator=g).cuda()
raw_f = torch.rand(n, sc + sf, 4, generator=g).cuda()
zf = (torch.sort(torch.rand(n, sc + sf, generator=g), -1).values * 4 + 2).cuda()
for _ in range(4):
    z, mids = ops.stratified_z(z_lin, t)
    ops.composite_forward(raw_c.reshape(-1, 4), z, d, True)
    ops.composite_forward(raw_f.reshape(-1, 4), zf, d, False)
torch.cuda.synchronize()
print("done")
