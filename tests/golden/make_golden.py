"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

It imports the reference's nerf/render.py, nerf/nerf.py, pi_GAN/render.py, pi_GAN/modules.py
(with empty stub modules for matplotlib / imageio / plyfile / skimage, which the path never
calls -- SURVEY.md 8c), runs them on CPU with fixed seeds and stores inputs + outputs of every
stage of the hot path.  Nothing from the reference is copied into the repo; only numbers are.
The GPU box has no /root/reference: tests read these .npz files instead.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import sys
import types

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


def sha_state(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def spy(mod, name, log):
    orig = getattr(mod, name)

    def wrapped(*a, **k):
        out = orig(*a, **k)
        log.append((name, a, out))
        return out
    setattr(mod, name, wrapped)
    return orig


def np_(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def sample_grad(g: torch.Tensor):
    g = g.detach().reshape(-1).double()
    return dict(sum=float(g.sum()), l2=float(g.norm()), sample=g[::97].float().numpy().copy())


def main():
    torch.set_num_threads(8)
    nr = load("ref_nerf_render", f"{REF}/nerf/render.py")
    nn_ = load("ref_nerf_nerf", f"{REF}/nerf/nerf.py")
    torch.autograd.set_detect_anomaly(False)
    for n in ["matplotlib", "matplotlib.pyplot", "imageio", "plyfile", "skimage", "skimage.measure"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, f"{REF}/pi_GAN")
    import render as pr          # noqa: E402  (pi_GAN/render.py)
    import modules as pm         # noqa: E402  (pi_GAN/modules.py)

    from msra_practice_project_b200 import models as my

    meta = {}

    # ---------------- weights: deterministic init equality ------------------------------------
    torch.manual_seed(0)
    ref_c, ref_f = nn_.NeRF(), nn_.NeRF()
    torch.manual_seed(0)
    my_c, my_f = my.NeRF(), my.NeRF()
    assert sha_state(ref_c.state_dict()) == sha_state(my_c.state_dict())
    assert sha_state(ref_f.state_dict()) == sha_state(my_f.state_dict())
    meta["nerf_seed0_coarse_sha"] = sha_state(ref_c.state_dict())
    meta["nerf_seed0_fine_sha"] = sha_state(ref_f.state_dict())

    torch.manual_seed(0)
    ref_film = pm.FilmSirenNeRF()
    torch.manual_seed(0)
    my_film = my.FilmSirenNeRF()
    assert sha_state(ref_film.state_dict()) == sha_state(my_film.state_dict())
    meta["film_seed0_sha"] = sha_state(ref_film.state_dict())

    # ---------------- linspace vectors -------------------------------------------------------
    lin = {}
    for (a, b, n) in [(2.0, 6.0, 64), (0.0, 1.0, 128), (0.0, 1.0, 64), (0.5, 1.5, 24), (0.0, 1.0, 24),
                      (2.0, 6.0, 16), (0.0, 1.0, 16), (0.5, 1.5, 8), (0.0, 1.0, 8), (0.0, 1.0, 33), (2.0, 6.0, 7)]:
        lin[f"lin_{a}_{b}_{n}"] = torch.linspace(a, b, steps=n).numpy()
    np.savez_compressed(f"{OUT}/linspace.npz", **lin)

    # ---------------- stand-alone kernels ----------------------------------------------------
    k = {}
    g = torch.Generator().manual_seed(1234)
    # get_rays, python-float focal (float32 result) and np.float64 focal (float64 result)
    pose = nr.__dict__.get("camera_pos_to_transform_matrix", None)
    c2w = pr.camera_pos_to_transform_matrix(4.0, 0.4, -0.5)
    o, d = nr.get_rays(20, 12, 20 * 1.3875, c2w)
    k["rays_c2w"] = c2w
    k["rays_o_20x12"] = np.ascontiguousarray(o); k["rays_d_20x12"] = d
    o64, d64 = nr.get_rays(9, 7, np.float64(9 / 2 / np.tan(6 * np.pi / 180)), c2w)
    k["rays_d_9x7_f64focal"] = d64
    # raw_to_outputs forward + autograd backward
    N, S = 37, 48
    raw = torch.rand(N, S, 4, generator=g)
    raw[..., 3] = torch.relu(torch.randn(N, S, generator=g)) * 3.0
    z = torch.sort(torch.rand(N, S, generator=g) * 4 + 2, -1).values
    dirs = torch.randn(N, 3, generator=g)
    raw.requires_grad_(True)
    rgb, depth, acc, w = nr.raw_to_outputs(raw, z, dirs)
    g_rgb = torch.randn(N, 3, generator=g); g_depth = torch.randn(N, generator=g); g_acc = torch.randn(N, generator=g)
    (d_raw,) = torch.autograd.grad((rgb * g_rgb).sum() + (depth * g_depth).sum() + (acc * g_acc).sum(), raw)
    k.update(c_raw=np_(raw), c_z=np_(z), c_dirs=np_(dirs), c_rgb=np_(rgb), c_depth=np_(depth), c_acc=np_(acc),
             c_w=np_(w), c_g_rgb=np_(g_rgb), c_g_depth=np_(g_depth), c_g_acc=np_(g_acc), c_d_raw=np_(d_raw))
    # sample_pdf: random, one-hot, all-zero, tiny weights; shared bins
    nb, Sf = 31, 40
    bins = torch.linspace(2.0, 6.0, nb + 1)
    bins = 0.5 * (bins[1:] + bins[:-1])
    ws = torch.rand(9, nb - 1, generator=g)
    ws[1] = 0; ws[1, 7] = 1.0                    # one-hot
    ws[2] = 0                                    # all zero -> uniform
    ws[3] = ws[3] * 1e-7                         # tiny
    ws[4, :10] = 0; ws[4, 20:] = 0               # empty ends
    sp = nr.sample_pdf(bins.expand(9, nb), ws, Sf)
    k.update(sp_bins=np_(bins), sp_w=np_(ws), sp_out=np_(sp), sp_u=torch.linspace(0., 1., steps=Sf).numpy())
    bins2 = torch.sort(torch.rand(5, 16, generator=g) * 3, -1).values    # per-ray bins
    ws2 = torch.rand(5, 15, generator=g)
    k.update(sp2_bins=np_(bins2), sp2_w=np_(ws2), sp2_out=np_(nr.sample_pdf(bins2, ws2, 33)), sp2_u=torch.linspace(0., 1., steps=33).numpy())
    # posenc + NeRF MLP + FiLM MLP on a few points
    x = torch.cat([torch.rand(96, 3, generator=g) * 8 - 4, torch.nn.functional.normalize(torch.randn(96, 3, generator=g), dim=-1)], -1)
    k["mlp_x"] = np_(x)
    k["posenc10"] = np_(ref_c.pe_pos(x[:, :3])); k["posenc4"] = np_(ref_c.pe_dir(x[:, 3:]))
    with torch.no_grad():
        k["nerf_seed0_coarse_out"] = np_(ref_c(x))
        k["nerf_seed0_fine_out"] = np_(ref_f(x))
    film = torch.cat([1.0 + 0.3 * torch.randn(9, 256, generator=g), 0.2 * torch.randn(9, 256, generator=g)], -1)
    xs = torch.cat([torch.rand(96, 3, generator=g) * 0.6 - 0.3, x[:, 3:]], -1)
    ref_film.set_film_params(film)
    with torch.no_grad():
        k["film_x"] = np_(xs); k["film_params"] = np_(film); k["film_seed0_out"] = np_(ref_film(xs))
    np.savez_compressed(f"{OUT}/kernels.npz", **k)

    # ---------------- staged NeRF render (C1 shape, reduced image) -----------------------------
    W = H = 12
    Sc = Sf = 64
    focal = W * 1.3875
    c2w = pr.camera_pos_to_transform_matrix(4.0, 0.3, -30 * np.pi / 180)
    log = []
    originals = {n: spy(nr, n, log) for n in ("run_network", "raw_to_outputs", "sample_pdf")}
    rays = np.reshape(np.transpose(np.stack(nr.get_rays(W, H, focal, c2w), 0), [1, 2, 0, 3]), [-1, 2, 3])
    torch.manual_seed(5)
    t_rand = torch.rand(W * H, Sc)             # same call/shape as nerf/render.py:131
    torch.manual_seed(5)
    with torch.no_grad():
        outs = nr.render_rays(torch.tensor(rays, dtype=torch.float), 2.0, 6.0, ref_c, ref_f, Sc, Sf)
    for n, f in originals.items():
        setattr(nr, n, f)
    st = dict(c2w=c2w, W=W, H=H, focal=focal, near=2.0, far=6.0, Sc=Sc, Sf=Sf, rays=rays.astype(np.float32), t_rand=np_(t_rand),
              z_lin=torch.linspace(2.0, 6.0, steps=Sc).numpy(), u=torch.linspace(0., 1., steps=Sf).numpy())
    (_, a0, raw_c), (_, a1, o1), (_, a2, zs), (_, a3, raw_f), (_, a4, o4) = log
    st.update(coarse_pts=np_(a0[0]), view_dirs=np_(a0[1]), raw_coarse=np_(raw_c), z_coarse=np_(a1[1]),
              weights_coarse=np_(o1[3]), mids=np_(a2[0][0]), z_samples=np_(zs), raw_fine=np_(raw_f),
              z_fine=np_(a4[1]), weights_fine=np_(o4[3]))
    for name, t in zip(["rgb_c", "depth_c", "acc_c", "rgb_f", "depth_f", "acc_f"], outs):
        st[name] = np_(t)
    np.savez_compressed(f"{OUT}/nerf_stages.npz", **st)

    # ---------------- NeRF train step grads (C3 shape, reduced) --------------------------------
    torch.manual_seed(0)
    tc, tf_ = nn_.NeRF(), nn_.NeRF()
    my.damp_nerf_(tc); my.damp_nerf_(tf_)
    nB, tSc, tSf = 24, 16, 16
    gg = torch.Generator().manual_seed(7)
    tr_rays = torch.tensor(rays[::6][:nB], dtype=torch.float)
    target = torch.rand(nB, 3, generator=gg); target_a = torch.rand(nB, generator=gg)
    torch.manual_seed(11)
    t_rand_tr = torch.rand(nB, tSc)
    torch.manual_seed(11)
    rc, _, ac, rf, _, af = nr.render_rays(tr_rays, 2.0, 6.0, tc, tf_, tSc, tSf)
    loss = ((rf - target) ** 2).mean() + ((rc - target) ** 2).mean() + 0.1 * ((ac - target_a) ** 2).mean() + 0.1 * ((af - target_a) ** 2).mean()
    loss.backward()
    tr = dict(rays=np_(tr_rays), target=np_(target), target_a=np_(target_a), t_rand=np_(t_rand_tr), loss=float(loss),
              rgb_c=np_(rc), rgb_f=np_(rf), acc_c=np_(ac), acc_f=np_(af), Sc=tSc, Sf=tSf,
              z_lin=torch.linspace(2.0, 6.0, steps=tSc).numpy(), u=torch.linspace(0., 1., steps=tSf).numpy())
    for tag, m in (("coarse", tc), ("fine", tf_)):
        for name, p in m.named_parameters():
            s = sample_grad(p.grad)
            tr[f"g_{tag}.{name}.sum"] = s["sum"]; tr[f"g_{tag}.{name}.l2"] = s["l2"]; tr[f"g_{tag}.{name}.sample"] = s["sample"]
    np.savez_compressed(f"{OUT}/nerf_train.npz", **tr)

    # ---------------- pi-GAN: staged render + grads wrt FiLM params + density grid ------------
    torch.manual_seed(0)
    film_net = pm.FilmSirenNeRF()
    gp = torch.Generator().manual_seed(3)
    film = torch.cat([1.0 + 0.2 * torch.randn(9, 256, generator=gp), 0.1 * torch.randn(9, 256, generator=gp)], -1).requires_grad_(True)
    film_net.set_film_params(film)
    pW = 8
    pfocal = pW / 2 / np.tan(12 / 2 * np.pi / 180)        # np.float64 as in pi_GAN/modules.py:127
    ppose = pr.camera_pos_to_transform_matrix(1, 0.2, -0.1)
    torch.manual_seed(21)
    pt_rand = torch.rand(pW * pW, 12)
    torch.manual_seed(21)
    img = pr.render_image(pW, pW, pfocal, ppose, 0.5, 1.5, film_net, film_net, 12, 12)
    gi = torch.randn(pW, pW, 3, generator=gp)
    (img * gi).sum().backward()
    pg = dict(film=np_(film), W=pW, focal=float(pfocal), pose=ppose, near=0.5, far=1.5, Sc=12, Sf=12, t_rand=np_(pt_rand),
              image=np_(img), g_image=np_(gi), g_film=np_(film.grad),
              z_lin=torch.linspace(0.5, 1.5, steps=12).numpy(), u=torch.linspace(0., 1., steps=12).numpy())
    for name, p in film_net.named_parameters():
        s = sample_grad(p.grad)
        pg[f"g.{name}.sum"] = s["sum"]; pg[f"g.{name}.l2"] = s["l2"]; pg[f"g.{name}.sample"] = s["sample"]
    # density grid (pi_GAN/utils.py:59-91 arithmetic, N=6)
    Ng = 6
    idx = torch.arange(0, Ng ** 3, 1, out=torch.LongTensor())
    smp = torch.zeros(Ng ** 3, 4)
    vs = 0.2 / (Ng - 1)
    smp[:, 2] = idx % Ng
    smp[:, 1] = torch.floor_divide(idx.long(), Ng) % Ng
    smp[:, 0] = torch.floor_divide(torch.floor_divide(idx.long(), Ng), Ng) % Ng
    smp[:, 0] = (smp[:, 0] * vs) + (-0.1); smp[:, 1] = (smp[:, 1] * vs) + (-0.1); smp[:, 2] = (smp[:, 2] * vs) + (-0.1)
    with torch.no_grad():
        sub = torch.cat([smp[:, :3], torch.zeros_like(smp[:, :3])], -1)
        pg["grid_N"] = Ng; pg["grid_pts"] = np_(smp[:, :3]); pg["grid_neg_sigma"] = np_(-film_net(sub)[:, 3])
    np.savez_compressed(f"{OUT}/pigan.npz", **pg)

    with open(f"{OUT}/meta.json", "w") as f:
        json.dump(dict(meta, torch=torch.__version__, numpy=np.__version__,
                       reference="JeffreyXiang/MSRA-practice-project @ /root/reference (read-only)"), f, indent=1)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
