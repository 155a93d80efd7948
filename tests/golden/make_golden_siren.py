"""Golden fixture for SirenNeRF from the UNMODIFIED reference (nerf/nerf.py:97-170).
Run in the build container only: python tests/golden/make_golden_siren.py -> tests/golden/siren.npz"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    nn_ = load("ref_nerf_nerf", f"{REF}/nerf/nerf.py")
    nr = load("ref_nerf_render", f"{REF}/nerf/render.py")
    torch.autograd.set_detect_anomaly(False)
    from msra_practice_project_b200 import models as my
    torch.manual_seed(0)
    ref = nn_.SirenNeRF()
    torch.manual_seed(0)
    mine = my.SirenNeRF()
    for (ka, a), (kb, b) in zip(ref.state_dict().items(), mine.state_dict().items()):
        assert ka == kb and torch.equal(a, b), ka
    g = torch.Generator().manual_seed(77)
    x = torch.cat([torch.rand(96, 3, generator=g) * 8 - 4, torch.nn.functional.normalize(torch.randn(96, 3, generator=g), dim=-1)], -1)
    out = {"x": x.numpy()}
    with torch.no_grad():
        out["out"] = ref(x).numpy()
    # a small training step through render_rays (coarse == fine model family): gradients of every parameter
    torch.manual_seed(0)
    c, f = nn_.SirenNeRF(), nn_.SirenNeRF()
    rays = torch.cat([torch.tensor([[0.0, 0.0, 4.0]]).expand(20, 3)[:, None], torch.nn.functional.normalize(
        torch.randn(20, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)[:, None]], 1)
    target = torch.rand(20, 3, generator=g)
    torch.manual_seed(3)
    t_rand = torch.rand(20, 16)
    torch.manual_seed(3)
    rc, _, _, rf, _, _ = nr.render_rays(rays, 2.0, 6.0, c, f, 16, 16)
    loss = ((rf - target) ** 2).mean() + ((rc - target) ** 2).mean()
    loss.backward()
    out.update(rays=rays.numpy(), target=target.numpy(), t_rand=t_rand.numpy(), rgb_c=rc.detach().numpy(), rgb_f=rf.detach().numpy(),
               loss=float(loss), z_lin=torch.linspace(2.0, 6.0, steps=16).numpy(), u=torch.linspace(0., 1., steps=16).numpy())
    for tag, m in (("coarse", c), ("fine", f)):
        for name, p in m.named_parameters():
            gr = p.grad.detach().reshape(-1).double()
            out[f"g_{tag}.{name}.l2"] = float(gr.norm())
            out[f"g_{tag}.{name}.sample"] = gr[::97].float().numpy().copy()
    np.savez_compressed(f"{OUT}/siren.npz", **out)
    print("wrote siren.npz", {k: np.asarray(v).shape for k, v in list(out.items())[:6]})


if __name__ == "__main__":
    main()
