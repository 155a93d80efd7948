"""CPU oracle for the volumetric ray-march render path (TEST INFRASTRUCTURE ONLY).

This file is a plain numpy (float32) restatement of the reference's algorithm for the hot
path.  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
package (``msra_practice_project_b200``) never imports anything from ``oracle/`` and has no
CPU fallback.

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build container
by ``tests/golden/make_golden.py`` (which imports ``/root/reference`` read-only) and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` replays them.

Every function cites the reference lines it restates (paths relative to /root/reference).
Arrays are float32 unless stated; accumulation order follows torch's CPU kernels where it
matters (sequential cumsum / cumprod along the sample axis).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# small host-side vectors
# --------------------------------------------------------------------------------------
def linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """``torch.linspace(start, end, steps)`` in float32 (nerf/render.py:35 ``u``, :123 coarse z).

    torch's CPU kernel fills SIMD-width chunks as ``base(chunk) + step*lane`` with a different
    base formula in each half of the vector, so its last-bit rounding depends on the build's
    vector width and cannot be restated portably.  The two vectors
    are therefore *inputs* of the path: the product makes them with ``torch.linspace`` on the
    host and so does the oracle.  Golden files store the vectors that were used.
    """
    import torch
    return torch.linspace(float(start), float(end), steps=int(steps), dtype=torch.float32).numpy().copy()


# --------------------------------------------------------------------------------------
# a1  get_rays                                                    nerf/render.py:7-23
# --------------------------------------------------------------------------------------
def get_rays(width: int, height: int, focal, c2w: np.ndarray):
    """Pinhole rays for every pixel (row-major, column index fastest).

    dirs = [(i - W/2)/f, -(j - H/2)/f, -1];  rays_d = R . dirs;  rays_o = c2w[:3, 3].
    """
    i, j = np.meshgrid(np.arange(width, dtype=F32), np.arange(height, dtype=F32), indexing="xy")
    dirs = np.stack([(i - width * 0.5) / focal, -(j - height * 0.5) / focal, -np.ones_like(i)], -1)
    rot = np.asarray(c2w)[:3, :3]
    rays_d = np.sum(dirs[..., None, :] * rot, -1)
    rays_o = np.broadcast_to(np.asarray(c2w)[:3, -1], rays_d.shape)
    return rays_o, rays_d


def image_rays(width, height, focal, c2w) -> np.ndarray:
    """The [H*W, 2, 3] float32 ray table render_image builds (nerf/render.py:151-154,159)."""
    o, d = get_rays(width, height, focal, c2w)
    rays = np.stack([o, d], 0)
    rays = np.transpose(rays, [1, 2, 0, 3]).reshape(-1, 2, 3)
    return rays.astype(F32)


# --------------------------------------------------------------------------------------
# a4 (first half)  stratified coarse z                           nerf/render.py:123-132
# --------------------------------------------------------------------------------------
def stratified_z(z_lin: np.ndarray, t_rand: np.ndarray):
    """z = lower + (upper - lower) * t_rand with mids of the *un-jittered* linspace.

    Returns (z_vals[N,Sc], mids[Sc-1]).  ``z_lin`` is linspace(near, far, Sc).
    """
    z_lin = np.asarray(z_lin, F32)
    mids = (F32(0.5) * (z_lin[1:] + z_lin[:-1])).astype(F32)
    upper = np.concatenate([mids, z_lin[-1:]])
    lower = np.concatenate([z_lin[:1], mids])
    z = (lower[None, :] + (upper - lower)[None, :] * np.asarray(t_rand, F32)).astype(F32)
    return z, mids


# --------------------------------------------------------------------------------------
# a6  positional encoding                                        nerf/nerf.py:44-49
# --------------------------------------------------------------------------------------
def posenc(x: np.ndarray, length: int) -> np.ndarray:
    """[sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] -- blocks of 3, no identity term."""
    x = np.asarray(x, F32)
    out = []
    for i in range(length):
        f = F32(2.0 ** i)
        out.append(np.sin(f * x, dtype=F32))
        out.append(np.cos(f * x, dtype=F32))
    return np.concatenate(out, -1).astype(F32)


def _linear(x, w, b):
    """nn.Linear: y = x W^T + b, weight [out, in] (nerf/nerf.py:21, pi_GAN/modules.py:23)."""
    return (x @ np.asarray(w, F32).T + np.asarray(b, F32)).astype(F32)


def _sigmoid(x):
    return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


# --------------------------------------------------------------------------------------
# a7  NeRF MLP                                                   nerf/nerf.py:75-94
# --------------------------------------------------------------------------------------
def nerf_mlp(p: dict, x: np.ndarray) -> np.ndarray:
    """x[M,6] = (pos, unit dir) -> [M,4] = (sigmoid rgb, relu sigma).

    ``p`` maps the reference state-dict keys (layers_pos.{0..7}, layers_dir.{0,1},
    output_layer_sigma, output_layer_rgb; .weight/.bias) to numpy arrays.
    """
    x = np.asarray(x, F32)
    pe = posenc(x[:, :3], 10)
    de = posenc(x[:, 3:], 4)
    relu = lambda a: np.maximum(a, F32(0))
    h = relu(_linear(pe, p["layers_pos.0.weight"], p["layers_pos.0.bias"]))
    for l in range(1, 5):
        h = relu(_linear(h, p[f"layers_pos.{l}.weight"], p[f"layers_pos.{l}.bias"]))
    h = np.concatenate([pe, h], -1)                                   # nerf.py:84 (pe first)
    h = relu(_linear(h, p["layers_pos.5.weight"], p["layers_pos.5.bias"]))
    h = relu(_linear(h, p["layers_pos.6.weight"], p["layers_pos.6.bias"]))
    h = relu(_linear(h, p["layers_pos.7.weight"], p["layers_pos.7.bias"]))
    sigma = relu(_linear(h, p["output_layer_sigma.weight"], p["output_layer_sigma.bias"]))
    h = _linear(h, p["layers_dir.0.weight"], p["layers_dir.0.bias"])   # linear, no activation
    h = np.concatenate([h, de], -1)                                   # nerf.py:90 (h first)
    h = relu(_linear(h, p["layers_dir.1.weight"], p["layers_dir.1.bias"]))
    rgb = _sigmoid(_linear(h, p["output_layer_rgb.weight"], p["output_layer_rgb.bias"]))
    return np.concatenate([rgb, sigma], -1).astype(F32)


# --------------------------------------------------------------------------------------
# (8f-1)  SirenNeRF MLP                                          nerf/nerf.py:97-170
# --------------------------------------------------------------------------------------
def siren_nerf_mlp(p: dict, x: np.ndarray) -> np.ndarray:
    """sin(30 (W x + b)) trunk on raw inputs, skip [pos | h] into layer 5, relu sigma head, linear layers_dir.0,
    sin layers_dir.1 on [h | dir], sigmoid rgb head.  Keys as NeRF."""
    x = np.asarray(x, F32)
    pos, d = x[:, :3], x[:, 3:]
    sin30 = lambda a: np.sin(F32(30.0) * a, dtype=F32)
    h = sin30(_linear(pos, p["layers_pos.0.weight"], p["layers_pos.0.bias"]))
    for l in range(1, 5):
        h = sin30(_linear(h, p[f"layers_pos.{l}.weight"], p[f"layers_pos.{l}.bias"]))
    h = np.concatenate([pos, h], -1)                                  # nerf.py:158 (pos first)
    for l in range(5, 8):
        h = sin30(_linear(h, p[f"layers_pos.{l}.weight"], p[f"layers_pos.{l}.bias"]))
    sigma = np.maximum(_linear(h, p["output_layer_sigma.weight"], p["output_layer_sigma.bias"]), F32(0))
    h = _linear(h, p["layers_dir.0.weight"], p["layers_dir.0.bias"])
    h = sin30(_linear(np.concatenate([h, d], -1), p["layers_dir.1.weight"], p["layers_dir.1.bias"]))
    rgb = _sigmoid(_linear(h, p["output_layer_rgb.weight"], p["output_layer_rgb.bias"]))
    return np.concatenate([rgb, sigma], -1).astype(F32)


# --------------------------------------------------------------------------------------
# a8  FiLM-SIREN MLP                          pi_GAN/modules.py:22-25, 96-99, 101-118
# --------------------------------------------------------------------------------------
def film_siren_mlp(p: dict, film: np.ndarray, x: np.ndarray, use_dir: bool = True, w0: float = 30.0,
                   return_sigma_only: bool = False) -> np.ndarray:
    """sin(w0 * (gamma * (W x + b) + beta)) layers; film[9,512] = gamma(256) || beta(256).

    Keys: input_layer, hidden_layers.{0..6}, output_layer_sigma.0, hidden_layer_rgb,
    output_layer_rgb.0 (.weight/.bias).
    """
    x = np.asarray(x, F32)
    film = np.asarray(film, F32)

    def fs(h, name, l):
        gamma, beta = film[l, :256], film[l, 256:]
        a = _linear(h, p[name + ".weight"], p[name + ".bias"])
        a = (gamma * a + beta).astype(F32)
        return np.sin(F32(w0) * a, dtype=F32)

    h = fs(x[:, :3], "input_layer", 0)
    for i in range(7):
        h = fs(h, f"hidden_layers.{i}", i + 1)
    sigma = np.maximum(_linear(h, p["output_layer_sigma.0.weight"], p["output_layer_sigma.0.bias"]), F32(0))
    if return_sigma_only:
        return sigma[:, 0]
    if use_dir:
        h = np.concatenate([h, x[:, 3:]], -1)
    h = fs(h, "hidden_layer_rgb", 8)
    rgb = _sigmoid(_linear(h, p["output_layer_rgb.0.weight"], p["output_layer_rgb.0.bias"]))
    return np.concatenate([rgb, sigma], -1).astype(F32)


# --------------------------------------------------------------------------------------
# a5  run_network                                                nerf/render.py:59-75
# --------------------------------------------------------------------------------------
def run_network(ray_samples: np.ndarray, view_dirs: np.ndarray, network, chunk: int = 1024 * 64):
    """Flatten points, broadcast the per-ray unit view dir, evaluate in chunks -> [N,S,4]."""
    n, s, _ = ray_samples.shape
    flat = ray_samples.reshape(-1, 3)
    vd = np.broadcast_to(view_dirs[:, None, :], ray_samples.shape).reshape(-1, 3)
    inputs = np.concatenate([flat, vd], -1).astype(F32)
    outs = [network(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)]
    return np.concatenate(outs).reshape(n, s, 4)


# --------------------------------------------------------------------------------------
# a9  raw_to_outputs (alpha compositing)                         nerf/render.py:78-103
# --------------------------------------------------------------------------------------
def raw_to_outputs(raw: np.ndarray, z_vals: np.ndarray, rays_d: np.ndarray):
    """dists=[dz..., 1e10]*|d|; alpha=1-exp(-sigma*dists); T=excl. cumprod(1-alpha+1e-10);
    w=alpha*T; rgb=sum w c + (1-acc) (white background always); depth=sum w z; acc=sum w."""
    raw = np.asarray(raw, F32)
    z = np.asarray(z_vals, F32)
    d = np.asarray(rays_d, F32)
    dists = z[:, 1:] - z[:, :-1]
    dists = np.concatenate([dists, np.full((z.shape[0], 1), 1e10, F32)], -1)
    norm = np.sqrt(np.sum(d * d, -1, keepdims=True, dtype=F32), dtype=F32)
    dists = (dists * norm).astype(F32)
    rgb = raw[..., :3]
    with np.errstate(over="ignore"):
        alpha = (F32(1.0) - np.exp(-raw[..., 3] * dists, dtype=F32)).astype(F32)
    q = np.concatenate([np.ones((z.shape[0], 1), F32), (F32(1.0) - alpha + F32(1e-10)).astype(F32)], -1)
    # torch's CPU cumprod/cumsum accumulate float32 inputs in double (at::acc_type<float,false>)
    # and round every prefix to float32
    trans = np.cumprod(q.astype(np.float64), -1).astype(F32)[:, :-1]
    weights = (alpha * trans).astype(F32)
    rgb_map = np.sum(weights[..., None] * rgb, -2, dtype=F32)
    depth_map = np.sum(weights * z, -1, dtype=F32)
    acc_map = np.sum(weights, -1, dtype=F32)
    rgb_map = (rgb_map + (F32(1.0) - acc_map[..., None])).astype(F32)
    return rgb_map, depth_map, acc_map, weights


def raw_to_outputs_backward(raw, z_vals, rays_d, g_rgb, g_depth, g_acc):
    """Analytic reverse-mode of raw_to_outputs wrt raw (float64 internally).

    The reference gets this from autograd (nerf/train_nerf.py:167); the closed form is
    SURVEY.md appendix A.3 and is pinned against torch autograd in the golden files.
    Returns d_raw[N,S,4].
    """
    raw = np.asarray(raw, np.float64)
    z = np.asarray(z_vals, np.float64)
    d = np.asarray(rays_d, np.float64)
    n, s = z.shape
    delta = np.concatenate([z[:, 1:] - z[:, :-1], np.full((n, 1), 1e10)], -1) * np.linalg.norm(d, axis=-1, keepdims=True)
    sigma = raw[..., 3]
    e = np.exp(-sigma * delta)
    alpha = 1.0 - e
    q = 1.0 - alpha + 1e-10
    trans = np.cumprod(np.concatenate([np.ones((n, 1)), q], -1), -1)[:, :-1]
    w = alpha * trans
    g_rgb = np.asarray(g_rgb, np.float64)
    g_depth = np.zeros(n) if g_depth is None else np.asarray(g_depth, np.float64)
    g_acc = np.zeros(n) if g_acc is None else np.asarray(g_acc, np.float64)
    gw = np.sum(g_rgb[:, None, :] * (raw[..., :3] - 1.0), -1) + g_acc[:, None] + g_depth[:, None] * z
    gc = w[..., None] * g_rgb[:, None, :]
    gww = gw * w
    suffix = np.cumsum(gww[:, ::-1], -1)[:, ::-1] - gww      # sum_{j>k} gw_j w_j
    g_alpha = gw * trans - suffix / q
    g_sigma = g_alpha * delta * e
    return np.concatenate([gc, g_sigma[..., None]], -1)


# --------------------------------------------------------------------------------------
# a10  sample_pdf                                                nerf/render.py:27-56
# --------------------------------------------------------------------------------------
def sample_pdf(bins: np.ndarray, weights: np.ndarray, n_samples: int, u: np.ndarray | None = None,
               return_cdf: bool = False):
    """Inverse-CDF resampling with deterministic u = linspace(0,1,n_samples).

    bins[N,nb] (or [nb]), weights[N,nb-1]. searchsorted(right=True) == count(cdf <= u).
    """
    weights = (np.asarray(weights, F32) + F32(1e-5)).astype(F32)
    n = weights.shape[0]
    bins = np.asarray(bins, F32)
    if bins.ndim == 1:
        bins = np.broadcast_to(bins[None, :], (n, bins.shape[0]))
    pdf = (weights / np.sum(weights, -1, keepdims=True, dtype=F32)).astype(F32)
    cdf = np.cumsum(pdf.astype(np.float64), -1).astype(F32)      # double accumulator, see raw_to_outputs
    cdf = np.concatenate([np.zeros((n, 1), F32), cdf], -1)
    if u is None:
        u = linspace_f32(0.0, 1.0, n_samples)
    u = np.asarray(u, F32)
    nb = cdf.shape[-1]
    inds = np.stack([np.searchsorted(cdf[r], u, side="right") for r in range(n)]).astype(np.int64)
    below = np.maximum(0, inds - 1)
    above = np.minimum(nb - 1, inds)
    cdf_b = np.take_along_axis(cdf, below, -1)
    cdf_a = np.take_along_axis(cdf, above, -1)
    bins_b = np.take_along_axis(bins, below, -1)
    bins_a = np.take_along_axis(bins, above, -1)
    denom = (cdf_a - cdf_b).astype(F32)
    denom = np.where(denom < F32(1e-5), F32(1.0), denom).astype(F32)
    t = ((u[None, :] - cdf_b) / denom).astype(F32)
    samples = (bins_b + t * (bins_a - bins_b)).astype(F32)
    if return_cdf:
        return samples, cdf, inds
    return samples


def sample_pdf_from_cdf(bins: np.ndarray, cdf: np.ndarray, u: np.ndarray):
    """The part of sample_pdf after the CDF (nerf/render.py:39-54): searchsorted(right=True),
    clamp, gather, denom<1e-5 -> 1, lerp.  Used for the "indices and samples bit-exact GIVEN identical
    CDFs" check (north_star): feed the CDF a kernel produced and compare the rest exactly.
    Returns (samples, inds)."""
    cdf = np.asarray(cdf, F32)
    n, nb = cdf.shape
    bins = np.asarray(bins, F32)
    if bins.ndim == 1:
        bins = np.broadcast_to(bins[None, :], (n, nb))
    u = np.asarray(u, F32)
    inds = np.stack([np.searchsorted(cdf[r], u, side="right") for r in range(n)]).astype(np.int64)
    below = np.maximum(0, inds - 1)
    above = np.minimum(nb - 1, inds)
    cdf_b = np.take_along_axis(cdf, below, -1); cdf_a = np.take_along_axis(cdf, above, -1)
    bins_b = np.take_along_axis(bins, below, -1); bins_a = np.take_along_axis(bins, above, -1)
    denom = (cdf_a - cdf_b).astype(F32)
    denom = np.where(denom < F32(1e-5), F32(1.0), denom).astype(F32)
    t = ((u[None, :] - cdf_b) / denom).astype(F32)
    return (bins_b + (t * (bins_a - bins_b)).astype(F32)).astype(F32), inds


def sample_pdf_tolerance(bins, weights, u, cdf_ulps: float = 8.0) -> np.ndarray:
    """Per-sample max-abs tolerance for comparing two evaluations of sample_pdf whose CDFs
    differ by a few float32 ulps (different summation order: torch CPU uses a double
    accumulator, CUB and warp scans do not).  The inverse CDF is piecewise linear with slope
    (bin width)/(cdf interval); in near-empty bins the interval is ~1e-5/sum(w) so a 6e-8
    CDF perturbation moves the sample by ~1e-2 bin widths.  The bound is
    cdf_ulps * 2^-24 * (largest slope among the hit interval and its neighbours) + 2e-6.
    Not part of the reference; used by tests only."""
    _, cdf, inds = sample_pdf(bins, weights, len(u), u=u, return_cdf=True)
    n, nb = cdf.shape
    bins = np.asarray(bins, F32)
    if bins.ndim == 1:
        bins = np.broadcast_to(bins[None, :], (n, nb))
    width = (bins[:, 1:] - bins[:, :-1]).astype(np.float64)
    den = np.maximum((cdf[:, 1:] - cdf[:, :-1]).astype(np.float64), 1e-5)
    slope = width / den                                                   # [n, nb-1]
    pad = np.pad(slope, ((0, 0), (1, 1)), mode="edge")
    slope3 = np.maximum(np.maximum(pad[:, :-2], pad[:, 1:-1]), pad[:, 2:])
    k = np.clip(inds - 1, 0, nb - 2)
    return (2e-6 + cdf_ulps * 2.0 ** -24 * np.take_along_axis(slope3, k, -1)).astype(np.float64)


# --------------------------------------------------------------------------------------
# a4  render_rays                                                nerf/render.py:106-147
# --------------------------------------------------------------------------------------
def render_rays(rays: np.ndarray, near: float, far: float, coarse_net, fine_net,
                coarse_sample_num: int, fine_sample_num: int, t_rand: np.ndarray,
                z_lin: np.ndarray | None = None, u: np.ndarray | None = None, stages: dict | None = None):
    """Coarse stratified pass -> sample_pdf on the un-jittered mids with weights[1:-1] ->
    sort(cat) -> fine net on all Sc+Sf samples.  ``t_rand`` replaces torch.rand (:131).
    ``coarse_net`` / ``fine_net`` are callables x[M,6] -> [M,4].  If ``stages`` is a dict the
    intermediate tensors are stored in it (same names as tests/golden/make_golden.py)."""
    rays = np.asarray(rays, F32)
    rays_o, rays_d = rays[:, 0], rays[:, 1]
    norm = np.sqrt(np.sum(rays_d * rays_d, -1, keepdims=True, dtype=F32), dtype=F32)
    view_dirs = (rays_d / norm).astype(F32)
    if z_lin is None:
        z_lin = linspace_f32(near, far, coarse_sample_num)
    z_vals, mids = stratified_z(z_lin, t_rand)
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]).astype(F32)
    raw_c = run_network(pts, view_dirs, coarse_net)
    rgb_c, depth_c, acc_c, weights = raw_to_outputs(raw_c, z_vals, rays_d)
    z_samples = sample_pdf(mids, weights[:, 1:-1], fine_sample_num, u=u)
    z_fine = np.sort(np.concatenate([z_vals, z_samples], -1), -1)
    pts_f = (rays_o[:, None, :] + rays_d[:, None, :] * z_fine[:, :, None]).astype(F32)
    raw_f = run_network(pts_f, view_dirs, fine_net)
    rgb_f, depth_f, acc_f, weights_f = raw_to_outputs(raw_f, z_fine, rays_d)
    if stages is not None:
        stages.update(z_coarse=z_vals, raw_coarse=raw_c, weights_coarse=weights, z_samples=z_samples,
                      z_fine=z_fine, raw_fine=raw_f, weights_fine=weights_f)
    return rgb_c, depth_c, acc_c, rgb_f, depth_f, acc_f


# --------------------------------------------------------------------------------------
# a2  render_image                                               nerf/render.py:150-167
# --------------------------------------------------------------------------------------
def render_image(width, height, focal, pose, near, far, coarse_net, fine_net, coarse_sample_num,
                 fine_sample_num, t_rand: np.ndarray, chunk: int = 1024 * 16):
    """Chunked render of every pixel; t_rand is [H*W, Sc] (the per-chunk torch.rand draws
    of the reference, concatenated)."""
    rays = image_rays(width, height, focal, pose)
    rgb, depth, acc = [], [], []
    for i in range(0, rays.shape[0], chunk):
        out = render_rays(rays[i:i + chunk], near, far, coarse_net, fine_net, coarse_sample_num,
                          fine_sample_num, t_rand[i:i + chunk])
        rgb.append(out[3]); depth.append(out[4]); acc.append(out[5])
    return (np.concatenate(rgb).reshape(height, width, 3), np.concatenate(depth).reshape(height, width, 1),
            np.concatenate(acc).reshape(height, width, 1))


# --------------------------------------------------------------------------------------
# a13  create_mesh density query                                 pi_GAN/utils.py:59-91
# --------------------------------------------------------------------------------------
def density_grid_points(n: int, begin: int = 0, count: int | None = None) -> np.ndarray:
    """xyz of grid samples [begin, begin+count): index -> (x=idx//N^2 %N, y=idx//N %N, z=idx%N),
    coordinate = index*voxel_size + origin (-0.1), voxel_size = 0.2/(N-1) (utils.py:56-72)."""
    count = n ** 3 - begin if count is None else count
    idx = np.arange(begin, begin + count, dtype=np.int64)
    voxel_size = 0.2 / (n - 1)
    out = np.zeros((count, 3), F32)
    out[:, 2] = (idx % n).astype(F32)
    out[:, 1] = ((idx // n) % n).astype(F32)
    out[:, 0] = ((idx // n // n) % n).astype(F32)
    # reference does float32 tensor * python float + python float -> float32 ops
    out[:, 0] = out[:, 0] * F32(voxel_size) + F32(-0.1)
    out[:, 1] = out[:, 1] * F32(voxel_size) + F32(-0.1)
    out[:, 2] = out[:, 2] * F32(voxel_size) + F32(-0.1)
    return out


def density_query(p: dict, film: np.ndarray, n: int, begin: int = 0, count: int | None = None) -> np.ndarray:
    """-sigma at grid points with zero view direction (utils.py:82-90)."""
    pts = density_grid_points(n, begin, count)
    x = np.concatenate([pts, np.zeros_like(pts)], -1)
    return -film_siren_mlp(p, film, x, return_sigma_only=True)


# --------------------------------------------------------------------------------------
# helpers shared by tests / bench (not part of the reference surface)
# --------------------------------------------------------------------------------------
def state_dict_to_numpy(sd) -> dict:
    return {k: np.ascontiguousarray(v.detach().cpu().numpy(), dtype=F32) for k, v in sd.items()}


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    """-10 log10(mse) (nerf/train_nerf.py:160, nerf/test_nerf.py:107)."""
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return float("inf") if mse == 0 else -10.0 * np.log10(mse)
