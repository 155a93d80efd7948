import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())   # run from the repository root
import numpy as np, torch
from msra_practice_project_b200 import models, ops, _lib, nerf_render, pigan_render
torch.manual_seed(0)
c, f = models.NeRF().cuda(), models.NeRF().cuda()
pose = pigan_render.camera_pos_to_transform_matrix(4.0, 0.3, -0.5)
for _ in range(2):
    with torch.no_grad():
        nerf_render.render_image_device(800, 800, 800 * 1.3875, pose, 2.0, 6.0, c, f, 64, 128)
torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
buf = (C.c_longlong * (148 * 16))()
print("rc", lib.b2r_dbg_prof(buf))
a = np.array(buf[:], dtype=np.float64).reshape(148, 16)
lead = a[0::2]
tot = lead[:, 4].mean()
print("INFERENCE fine pass: mma total cycles %.0f  wait_wfull %.1f%%  wait_act %.1f%%" % (tot, 100 * lead[:, 2].mean() / tot, 100 * lead[:, 3].mean() / tot))
print("producer wait_empty %.1f%% of %.0f" % (100 * a[:, 0].mean() / a[:, 1].mean(), a[:, 1].mean()))
print("epi wait_acc %.1f%%" % (100 * a[:, 11].mean() / a[:, 13].mean()))
