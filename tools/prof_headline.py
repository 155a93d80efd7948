"""ncu target for the headline shape: ONE 800x800, 64+128 frame = the 40.96 M-row coarse launch and the 122.88 M-row fine launch
of nerf_tc_kernel (plus the small stages).  Usage on a B200:
    python tools/prof_headline.py && ncu --set full --clock-control none --import-source on -k regex:nerf_tc_kernel -c 2 \
        -f -o gpurun_out/r2_headline python tools/prof_headline.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msra_practice_project_b200 import models, nerf_render, pigan_render  # noqa: E402

torch.cuda.set_device(0)
torch.manual_seed(0)
coarse, fine = models.NeRF().cuda(), models.NeRF().cuda()
pose = pigan_render.camera_pos_to_transform_matrix(4.0, 0.3, -30 * np.pi / 180)
torch.manual_seed(5)
t = torch.rand(640000, 64, device="cuda")
with torch.no_grad():
    out = nerf_render.render_image_device(800, 800, 800 * 1.3875, pose, 2.0, 6.0, coarse, fine, 64, 128, t_rand=t)
torch.cuda.synchronize()
print("frame rendered: rgb mean %.5f" % float(out[3].mean()))
