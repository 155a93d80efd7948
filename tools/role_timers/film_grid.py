# Role timers of film_tc_kernel<false> on the 256^3 sigma-only grid query (needs the instrumented build: see README.md)
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())   # run from the repository root
import numpy as np, torch
from msra_practice_project_b200 import models, _lib, pigan_render
torch.manual_seed(0)
dev = "cuda"
net = models.FilmSirenNeRF().to(dev)
g = torch.Generator().manual_seed(0)
net.set_film_params(torch.cat([1.0 + 0.2 * torch.randn(9, 256, generator=g), 0.1 * torch.randn(9, 256, generator=g)], -1).to(dev))
n = 256
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pigan_render.density_grid(net, n, max_batch=n ** 3)
    e1.record()
    torch.cuda.synchronize()
print("grid ms", e0.elapsed_time(e1))
lib = C.CDLL(_lib.LIB_PATH)
if not hasattr(lib, "b2r_dbg_prof"):
    sys.exit(0)          # shipped (un-instrumented) library: timing only
buf = (C.c_longlong * (148 * 16))()
print("rc", lib.b2r_dbg_prof(buf))
a = np.array(buf[:], dtype=np.float64).reshape(148, 16)
lead = a[0::2]
tot = lead[:, 4].mean()
print("FiLM grid: mma loop cycles %.0f  wait_wfull %.1f%%  wait_act %.1f%%" % (tot, 100 * lead[:, 2].mean() / tot, 100 * lead[:, 3].mean() / tot))
print("producer wait_empty %.1f%% of %.0f" % (100 * a[:, 0].mean() / a[:, 1].mean(), a[:, 1].mean()))
print("epilogue warp 0: wait_acc %.1f%%, input stages %.1f%% of %.0f" % (100 * a[:, 11].mean() / a[:, 13].mean(), 100 * a[:, 12].mean() / a[:, 13].mean(), a[:, 13].mean()))
