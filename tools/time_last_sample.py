"""Cost breakdown of the last-sample sign check on the headline frame (800x800, 64+128): fused kernel with / without the
flagging code active, the fp32 re-evaluation alone, and the whole frame with the check on / off."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msra_practice_project_b200 import _lib, models, nerf_render, ops, pigan_render  # noqa: E402


def ev_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    W = H = 800
    torch.manual_seed(0)
    coarse, fine = models.NeRF().to(dev), models.NeRF().to(dev)
    pose = pigan_render.camera_pos_to_transform_matrix(4.0, 0.3, -30 * np.pi / 180)
    rays = ops.raygen(W, H, W * 1.3875, pose)
    torch.manual_seed(5)
    t = torch.rand(W * H, 64, device=dev)
    st = {}
    with torch.no_grad():
        nerf_render.render_rays(rays, 2.0, 6.0, coarse, fine, 64, 128, t_rand=t, stages=st)
    lib = _lib.lib()
    s = torch.cuda.current_stream(dev).cuda_stream
    for name, model, z in (("coarse", coarse, st["z_coarse"]), ("fine", fine, st["z_fine"])):
        flat = models.flat_params(model).detach()
        packed = ops._packed_weights(model, 0, flat, None, True)
        inp, rows, keep = ops._make_input(rays, z, None, None)
        raw = torch.empty((rows, 4), device=dev)
        spr = z.shape[1]
        t_off = ev_time(lambda: _lib.check(lib.b2r_mlp_tc_fwd(0, packed.data_ptr(), 1, C.byref(inp), raw.data_ptr(), 0, None, s), "tc"))
        for rel in (2.0 ** -8, 2.0 ** -7):
            ops.set_last_sample_band(0, rel)
            ls, k2 = ops._last_sample_begin(0, rows, spr, dev)

            def with_flags():
                k2[0].zero_()
                _lib.check(lib.b2r_mlp_tc_fwd(0, packed.data_ptr(), 1, C.byref(inp), raw.data_ptr(), 0, C.byref(ls), s), "tc")
            t_on = ev_time(with_flags)
            n = int(k2[0].item())
            ws_bytes = lib.b2r_mlp_f32_workspace_bytes(0, n, 0) + ((n * 4 + 15) & ~15)
            ws = torch.empty((ws_bytes // 4,), device=dev)
            t_fix = ev_time(lambda: _lib.check(lib.b2r_mlp_f32_last_sigma(0, flat.data_ptr(), None, 1, 1, 0, C.byref(inp), spr, k2[1].data_ptr(), n,
                                                                           raw.data_ptr(), ws.data_ptr(), ws_bytes, s), "fix"), reps=10)
            print(f"{name}: rows {rows}, kernel without flags {t_off:.3f} ms, with flags (rel {rel:.5f}) {t_on:.3f} ms, flagged {n} "
                  f"({100.0 * n / z.shape[0]:.2f} %), fp32 re-evaluation {t_fix:.3f} ms = {n * 979456e-9 / t_fix:.1f} TFLOP/s")
    ops.set_last_sample_band(0, 2.0 ** -7)
    for on in (False, True):
        def frame():
            with torch.no_grad():
                nerf_render.render_image_device(W, H, W * 1.3875, pose, 2.0, 6.0, coarse, fine, 64, 128, t_rand=t, exact_last_sample=on)
        frame(); frame()
        print(f"whole frame, check {'on ' if on else 'off'}: {ev_time(frame):.3f} ms")


if __name__ == "__main__":
    main()
