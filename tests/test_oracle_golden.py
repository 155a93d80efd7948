"""Pin the numpy oracle against outputs of the unmodified reference (tests/golden/*.npz,
made by tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import render_oracle as orc
from msra_practice_project_b200 import models

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def seeded_nerf():
    torch.manual_seed(0)
    return models.NeRF(), models.NeRF()


def test_weights_match_reference_init():
    import hashlib
    meta = json.load(open(os.path.join(GOLDEN, "meta.json")))

    def sha(sd):
        h = hashlib.sha256()
        for k in sd:
            h.update(k.encode()); h.update(np.ascontiguousarray(sd[k].numpy()).tobytes())
        return h.hexdigest()
    c, f = seeded_nerf()
    assert sha(c.state_dict()) == meta["nerf_seed0_coarse_sha"]
    assert sha(f.state_dict()) == meta["nerf_seed0_fine_sha"]
    torch.manual_seed(0)
    assert sha(models.FilmSirenNeRF().state_dict()) == meta["film_seed0_sha"]
    assert models.NERF_NUMEL == 593924 and models.FILM_NUMEL == 529156


def test_linspace_bit_exact(golden):
    for key, ref in golden.linspace.items():
        _, a, b, n = key.split("_")
        got = orc.linspace_f32(float(a), float(b), int(n))
        # torch.linspace's last bit depends on the CPU's SIMD width (see oracle docstring):
        # 1-ulp agreement here; every stage test feeds the vectors stored in the golden files.
        np.testing.assert_allclose(got, ref, rtol=2.4e-7, atol=0, err_msg=key)


def test_get_rays(golden):
    k = golden.kernels
    o, d = orc.get_rays(20, 12, 20 * 1.3875, k["rays_c2w"])
    assert d.dtype == np.float32
    assert np.array_equal(d, k["rays_d_20x12"]) and np.array_equal(o, k["rays_o_20x12"])
    _, d64 = orc.get_rays(9, 7, np.float64(9 / 2 / np.tan(6 * np.pi / 180)), k["rays_c2w"])
    assert np.array_equal(d64, k["rays_d_9x7_f64focal"])


def test_composite_forward_and_backward(golden):
    k = golden.kernels
    rgb, depth, acc, w = orc.raw_to_outputs(k["c_raw"], k["c_z"], k["c_dirs"])
    np.testing.assert_allclose(w, k["c_w"], atol=2e-7, rtol=0)
    np.testing.assert_allclose(rgb, k["c_rgb"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(depth, k["c_depth"], atol=4e-6, rtol=0)
    np.testing.assert_allclose(acc, k["c_acc"], atol=1e-6, rtol=0)
    d_raw = orc.raw_to_outputs_backward(k["c_raw"], k["c_z"], k["c_dirs"], k["c_g_rgb"], k["c_g_depth"], k["c_g_acc"])
    ref = k["c_d_raw"].astype(np.float64)
    # last sample: delta = 1e10|d| so d sigma is 0 (sigma>0) or O(1e10) (sigma==0); compare relatively
    scale = np.maximum(1.0, np.abs(ref))
    assert np.max(np.abs(d_raw - ref) / scale) < 2e-4


def test_sample_pdf(golden):
    k = golden.kernels
    out = orc.sample_pdf(k["sp_bins"], k["sp_w"], k["sp_out"].shape[1], u=k["sp_u"])
    np.testing.assert_allclose(out, k["sp_out"], atol=1e-6, rtol=0)
    out2 = orc.sample_pdf(k["sp2_bins"], k["sp2_w"], k["sp2_out"].shape[1], u=k["sp2_u"])
    np.testing.assert_allclose(out2, k["sp2_out"], atol=4e-6, rtol=0)   # values up to 3: a few ulp
    # KATs: uniform weights + uniform bins -> linear; u=0 -> bins[0]
    bins = np.linspace(2, 6, 33, dtype=np.float32)
    z = orc.sample_pdf(bins, np.ones((1, 32), np.float32), 17)
    np.testing.assert_allclose(z[0], bins[0] + orc.linspace_f32(0, 1, 17) * (bins[-1] - bins[0]), atol=2e-6)
    assert z[0, 0] == bins[0]


def test_posenc_and_mlps(golden):
    k = golden.kernels
    x = k["mlp_x"]
    np.testing.assert_allclose(orc.posenc(x[:, :3], 10), k["posenc10"], atol=2e-4, rtol=0)  # 2^9*x: 1ulp of arg
    np.testing.assert_allclose(orc.posenc(x[:, 3:], 4), k["posenc4"], atol=1e-6, rtol=0)
    c, f = seeded_nerf()
    pc, pf = orc.state_dict_to_numpy(c.state_dict()), orc.state_dict_to_numpy(f.state_dict())
    np.testing.assert_allclose(orc.nerf_mlp(pc, x), k["nerf_seed0_coarse_out"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(orc.nerf_mlp(pf, x), k["nerf_seed0_fine_out"], atol=2e-4, rtol=0)
    torch.manual_seed(0)
    fm = models.FilmSirenNeRF()
    out = orc.film_siren_mlp(orc.state_dict_to_numpy(fm.state_dict()), k["film_params"], k["film_x"])
    np.testing.assert_allclose(out, k["film_seed0_out"], atol=2e-4, rtol=0)


def test_nerf_stages_teacher_forced(golden):
    """Each stage of render_rays fed with the reference's own inputs for that stage."""
    s = golden.nerf_stages
    z, mids = orc.stratified_z(s["z_lin"], s["t_rand"])
    assert np.array_equal(z, s["z_coarse"]) and np.array_equal(mids, s["mids"])
    rays = s["rays"]
    pts = rays[:, None, 0, :] + rays[:, None, 1, :] * z[:, :, None]
    assert np.array_equal(pts.astype(np.float32), s["coarse_pts"])
    rgb, depth, acc, w = orc.raw_to_outputs(s["raw_coarse"], s["z_coarse"], rays[:, 1])
    np.testing.assert_allclose(w, s["weights_coarse"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(rgb, s["rgb_c"], atol=2e-6, rtol=0)
    zs = orc.sample_pdf(s["mids"], s["weights_coarse"][:, 1:-1], int(s["Sf"]), u=s["u"])
    tol = orc.sample_pdf_tolerance(s["mids"], s["weights_coarse"][:, 1:-1], s["u"])
    err = np.abs(zs.astype(np.float64) - s["z_samples"])
    assert np.all(err <= tol), float(np.max(err / tol))
    assert np.mean(err <= 2e-6) > 0.98          # the ill-conditioned (near-empty-bin) samples are rare
    zf = np.sort(np.concatenate([s["z_coarse"], s["z_samples"]], -1), -1)
    assert np.array_equal(zf, s["z_fine"])
    rgb, depth, acc, w = orc.raw_to_outputs(s["raw_fine"], s["z_fine"], rays[:, 1])
    np.testing.assert_allclose(rgb, s["rgb_f"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(depth, s["depth_f"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(acc, s["acc_f"], atol=2e-6, rtol=0)
    c, _ = seeded_nerf()
    pc = orc.state_dict_to_numpy(c.state_dict())
    raw = orc.run_network(s["coarse_pts"][:16], s["view_dirs"][:16], lambda x: orc.nerf_mlp(pc, x))
    np.testing.assert_allclose(raw, s["raw_coarse"][:16], atol=3e-4, rtol=0)


def test_density_grid(golden):
    p = golden.pigan
    n = int(p["grid_N"])
    assert np.array_equal(orc.density_grid_points(n), p["grid_pts"])
    torch.manual_seed(0)
    fm = models.FilmSirenNeRF()
    out = orc.density_query(orc.state_dict_to_numpy(fm.state_dict()), p["film"], n)
    np.testing.assert_allclose(out, p["grid_neg_sigma"], atol=2e-4, rtol=0)


def test_torch_port_matches_reference_goldens(golden):
    """oracle/torch_port.py (the CPU-baseline port timed by bench.py) reproduces the reference's outputs."""
    from oracle import torch_port as tp
    s = golden.nerf_stages
    c, f = seeded_nerf()
    sc_, sf_ = dict(c.state_dict()), dict(f.state_dict())
    with torch.no_grad():
        out = tp.render_rays(torch.from_numpy(s["rays"]), 2.0, 6.0, lambda x: tp.nerf_mlp(sc_, x), lambda x: tp.nerf_mlp(sf_, x),
                             64, 64, torch.from_numpy(s["t_rand"]))
    for got, name in zip(out, ["rgb_c", "depth_c", "acc_c", "rgb_f", "depth_f", "acc_f"]):
        np.testing.assert_allclose(got.numpy(), s[name], atol=1e-6, rtol=0, err_msg=name)


def test_siren_nerf_oracle_and_init(golden):
    """SirenNeRF (nerf/nerf.py:97-170): same init stream as the reference (checked when the fixture was
    made, tests/golden/make_golden_siren.py) and the numpy restatement against the reference's forward."""
    g = golden.siren
    torch.manual_seed(0)
    m = models.SirenNeRF()
    assert models.model_kind(m) == models.KIND_SIREN and sum(p.numel() for p in m.parameters()) == models.SIREN_NUMEL == 562052
    p = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    out = orc.siren_nerf_mlp(p, g["x"])
    # 30x sine argument gain per layer: fp32 round-off differences between numpy and ATen matmuls reach a few 1e-5
    np.testing.assert_allclose(out[:, :3], g["out"][:, :3], atol=2e-4, rtol=0)
    np.testing.assert_allclose(out[:, 3], g["out"][:, 3], atol=2e-4, rtol=1e-3)


def test_generator_host_side_matches_reference(golden):
    """models.Generator / MappingNetwork / Renderer (pi_GAN/modules.py:34-68,120-197): same state-dict keys and seed-0 init as
    the reference (SHA), the mapping network's film_params[B,9,512] on the golden z, and the (theta, phi) draws in the
    reference's order -- the host side of Generator.forward; its render is a -m gpu test."""
    import hashlib
    gg = golden.generator
    torch.manual_seed(0)
    gen = models.Generator(256, 8, near=0.5, far=1.5, fov=12, coarse_samples=8, fine_samples=8)
    h = hashlib.sha256()
    sd = gen.state_dict()
    for k in sd:
        h.update(k.encode()); h.update(np.ascontiguousarray(sd[k].numpy()).tobytes())
    assert h.hexdigest() == str(gg["state_sha"])
    with torch.no_grad():
        film = gen.get_mapping(torch.from_numpy(gg["z"]))
    assert tuple(film.shape) == (2, 9, 512)
    np.testing.assert_allclose(film.numpy(), gg["film"], atol=2e-6, rtol=0)
    assert np.all(film.numpy()[:, :, :256].mean(-1) > 0.9)          # gamma ~ 1, beta ~ 0: the head-bias init (modules.py:56-58)
    np.random.seed(3)
    r = gen.renderer
    from msra_practice_project_b200 import pigan_render
    for i in range(2):
        pose = r.draw_pose()
        want = pigan_render.camera_pos_to_transform_matrix(1, gg["theta_phi"][i, 0], gg["theta_phi"][i, 1])
        np.testing.assert_array_equal(pose, want)
    gen.set_resolution(16)
    assert r.width == 16 and r.height == 16 and abs(r.focal - 16 / 2 / np.tan(6 * np.pi / 180)) < 1e-9
    with pytest.raises(RuntimeError):
        gen(torch.from_numpy(gg["z"]))                                # CPU model: the render path has no CPU fallback
