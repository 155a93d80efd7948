"""Golden fixture for the pi-GAN Generator from the UNMODIFIED reference (pi_GAN/modules.py:34-68,120-197).
Run in the build container only: python tests/golden/make_golden_generator.py -> tests/golden/generator.npz"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))


def sha_state(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def main():
    for n in ["matplotlib", "matplotlib.pyplot", "imageio", "plyfile", "skimage", "skimage.measure"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, f"{REF}/pi_GAN")
    import modules as pm         # noqa: E402  (pi_GAN/modules.py; star-imports pi_GAN/render.py)
    from msra_practice_project_b200 import models as my

    kw = dict(near=0.5, far=1.5, fov=12, coarse_samples=8, fine_samples=8)
    torch.manual_seed(0)
    ref = pm.Generator(256, 8, **kw)
    torch.manual_seed(0)
    mine = my.Generator(256, 8, **kw)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    assert sha_state(ref.state_dict()) == sha_state(mine.state_dict())
    g = torch.Generator().manual_seed(11)
    z = torch.randn(2, 256, generator=g)
    out = {"z": z.numpy(), "state_sha": sha_state(ref.state_dict())}
    with torch.no_grad():
        out["film"] = ref.get_mapping(z).numpy()
        assert torch.allclose(mine.get_mapping(z), ref.get_mapping(z), atol=1e-6)
    # poses: Renderer.__call__ draws theta then phi per latent from np.random
    np.random.seed(3)
    draws = np.random.randn(4)
    out["theta_phi"] = np.stack([draws[0::2] * 0.3, draws[1::2] * 0.15], -1)
    # forward + backward of an image loss, CPU fp32; jitter = the torch.rand stream after manual_seed(5)
    np.random.seed(3)
    torch.manual_seed(5)
    img = ref(z)                                           # [2,3,8,8]
    target = torch.rand(2, 3, 8, 8, generator=g)
    loss = ((img - target) ** 2).mean()
    loss.backward()
    torch.manual_seed(5)
    out["t_rand"] = torch.stack([torch.rand(64, 8) for _ in range(2)]).numpy()
    out.update(img=img.detach().numpy(), target=target.numpy(), loss=float(loss))
    for name, p in ref.named_parameters():
        gr = p.grad.detach().reshape(-1).double()
        out[f"g.{name}.l2"] = float(gr.norm())
        out[f"g.{name}.sample"] = gr[::53].float().numpy().copy()
    np.savez_compressed(f"{OUT}/generator.npz", **out)
    print("wrote generator.npz; loss", float(loss), "img mean", float(img.mean()))


if __name__ == "__main__":
    main()
