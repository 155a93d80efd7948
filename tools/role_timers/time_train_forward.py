import sys, os
sys.path.insert(0, os.getcwd())   # run from the repository root
import numpy as np, torch
from msra_practice_project_b200 import models, ops
torch.manual_seed(0)
f = models.NeRF().cuda()
rays = ops.raygen(64, 64, 64 * 1.3875, np.eye(4)[:3])
z = torch.sort(torch.rand(4096, 192, device="cuda") * 4 + 2, -1).values.contiguous()
flat = models.flat_params(f).detach()
packed = ops.pack_tc(flat, 0)
for _ in range(3):
    raw, saved = ops.tc_train_forward(packed, 0, rays, z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    raw, saved = ops.tc_train_forward(packed, 0, rays, z)
e1.record(); torch.cuda.synchronize()
print("train fwd fine pass (786432 rows): %.3f ms" % (e0.elapsed_time(e1) / 10))
with torch.no_grad():
    for _ in range(3): r2 = ops.mlp(f, rays=rays, z=z, exact_last_sample=False)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): r2 = ops.mlp(f, rays=rays, z=z, exact_last_sample=False)
    e1.record(); torch.cuda.synchronize()
print("inference same rows: %.3f ms" % (e0.elapsed_time(e1) / 10))
