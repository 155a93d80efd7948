import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())   # run from the repository root
import numpy as np, torch
from msra_practice_project_b200 import models, ops, _lib
from msra_practice_project_b200.train_step import NerfTrainStep
torch.manual_seed(0)
c, f = models.NeRF().cuda(), models.NeRF().cuda()
step = NerfTrainStep(c, f, 2.0, 6.0, 64, 128, 4096, graph=False)
rays = ops.raygen(64, 64, 64 * 1.3875, np.eye(4)[:3])
rgb = torch.rand(4096, 3, device="cuda")
for _ in range(3):
    step(rays, rgb)
torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
buf = (C.c_longlong * (148 * 16))()
print("rc", lib.b2r_dbg_prof(buf))
a = np.array(buf[:], dtype=np.float64).reshape(148, 16)
names = ["prod.wait_empty", "prod.total", "mma.wait_wfull", "mma.wait_act", "mma.total", "spill0.wait_ready", "spill0.wait_read", "spill0.tail",
         "spill1.wait_ready", "spill1.wait_read", "spill1.tail", "epi.wait_acc", "epi.wait_spill", "epi.total"]
lead = a[0::2]; peer = a[1::2]
for i, n in enumerate(names):
    print(f"{n:20s} leader mean {lead[:, i].mean():12.0f}  peer mean {peer[:, i].mean():12.0f}   (cycles)")
tot = lead[:, 4].mean()
print("mma total cycles", tot, " wait_wfull %.1f%%  wait_act %.1f%%" % (100 * lead[:, 2].mean() / tot, 100 * lead[:, 3].mean() / tot))
et = a[:, 13].mean()
print("epi total", et, " wait_acc %.1f%% wait_spill %.1f%%" % (100 * a[:, 11].mean() / et, 100 * a[:, 12].mean() / et))
st = a[:, 1].mean()
print("spill: wait_ready %.1f%% wait_read %.1f%% of producer total" % (100 * a[:, 5].mean() / st, 100 * a[:, 6].mean() / st))
print("producer wait_empty %.1f%%" % (100 * a[:, 0].mean() / st))
