"""Per-parameter comparison of the bf16 tensor-core training path with (a) the fp32 layer-wise path and (b) a torch
emulation of the same bf16 pipeline (debug aid)."""
import sys
import torch
from msra_practice_project_b200 import models, ops

n, s = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (700, 64)
g = torch.Generator().manual_seed(1)
torch.manual_seed(0)
net = models.damp_nerf_(models.NeRF()).cuda()
o = torch.tensor([0.0, 0.0, 4.0]).expand(n, 3)
d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.25 + torch.tensor([0.0, 0.0, -1.0]), dim=-1) * 1.1
rays = torch.stack([o, d], 1).cuda()
z = (torch.sort(torch.rand(n, s, generator=g), -1).values * 4 + 2).cuda()
up = torch.randn(n * s, 4, generator=g).cuda()
res = {}
for mode in ("fp32", "bf16"):
    ops.set_grad_precision(mode)
    net.zero_grad(set_to_none=True)
    raw = ops.mlp(net, rays=rays, z=z)
    (raw * up).sum().backward()
    torch.cuda.synchronize()
    res[mode] = (raw.detach().clone(), {k: p.grad.detach().clone() for k, p in net.named_parameters()})


def emulate(net, rays, z, up):
    """the kernels' arithmetic in torch: bf16-rounded operands, fp32 accumulation, bf16-rounded saved tensors"""
    bf = lambda t: t.to(torch.bfloat16).float()
    P = {k: v.detach() for k, v in net.named_parameters()}
    W = lambda k: bf(P[k + ".weight"])
    B = lambda k: P[k + ".bias"]
    pts = (rays[:, None, 0] + rays[:, None, 1] * z[..., None]).reshape(-1, 3)
    vd = torch.nn.functional.normalize(rays[:, 1], dim=-1)[:, None].expand(-1, z.shape[1], -1).reshape(-1, 3)
    enc = lambda x, L: torch.cat([f(x * 2.0 ** i) for i in range(L) for f in (torch.sin, torch.cos)], -1)
    pe, de = bf(enc(pts, 10)), bf(enc(vd, 4))
    h, hs = pe, []
    for l in range(8):
        x = torch.cat([pe, h], -1) if l == 5 else h
        hs.append(x)
        pre = x @ W(f"layers_pos.{l}").T + B(f"layers_pos.{l}")
        h32 = torch.relu(pre)
        h = bf(h32)
    sig_pre = h32 @ P["output_layer_sigma.weight"].T + P["output_layer_sigma.bias"]
    gl = bf(h @ W("layers_dir.0").T + B("layers_dir.0"))
    xd = torch.cat([gl, de], -1)
    hd32 = torch.relu(xd @ W("layers_dir.1").T + B("layers_dir.1"))
    hd = bf(hd32)
    rgb = torch.sigmoid(hd32 @ P["output_layer_rgb.weight"].T + P["output_layer_rgb.bias"])
    raw = torch.cat([rgb, torch.relu(sig_pre)], -1)
    G = {}
    gc = up[:, :3] * rgb * (1 - rgb)
    gs = up[:, 3:] * (sig_pre > 0)
    G["output_layer_rgb.weight"], G["output_layer_rgb.bias"] = gc.T @ hd, gc.sum(0)
    G["output_layer_sigma.weight"], G["output_layer_sigma.bias"] = gs.T @ h, gs.sum(0)
    gd1 = bf((gc @ P["output_layer_rgb.weight"]) * (hd > 0))
    G["layers_dir.1.weight"], G["layers_dir.1.bias"] = gd1.T @ xd, gd1.sum(0)
    gg = bf(gd1 @ W("layers_dir.1")[:, :256])
    G["layers_dir.0.weight"], G["layers_dir.0.bias"] = gg.T @ h, gg.sum(0)
    gh = bf((gg @ W("layers_dir.0") + gs * P["output_layer_sigma.weight"]) * (h > 0))
    for l in range(7, -1, -1):
        G[f"layers_pos.{l}.weight"], G[f"layers_pos.{l}.bias"] = gh.T @ hs[l], gh.sum(0)
        if l == 0:
            break
        w = W(f"layers_pos.{l}")
        if l == 5:
            w = w[:, 60:]
        prev = hs[l][:, 60:] if l == 5 else hs[l]
        gh = bf((gh @ w) * (prev > 0))
    return raw, G


torch.backends.cuda.matmul.allow_tf32 = False
raw_e, G_e = emulate(net, rays, z, up)
a, b = res["fp32"][0], res["bf16"][0]
print("raw max-abs diff vs fp32: rgb %.4g sigma %.4g (sigma max %.3g)" % ((a[:, :3] - b[:, :3]).abs().max(), (a[:, 3] - b[:, 3]).abs().max(), a[:, 3].max()))
print("raw max-abs diff vs emulation: rgb %.4g sigma %.4g" % ((raw_e[:, :3] - b[:, :3]).abs().max(), (raw_e[:, 3] - b[:, 3]).abs().max()))
for k in res["fp32"][1]:
    ga, gb, ge = res["fp32"][1][k], res["bf16"][1][k], G_e[k].reshape(res["bf16"][1][k].shape)
    print("%-28s |fp32| %.4e |bf16| %.4e rel-vs-fp32 %.4g rel-vs-emulation %.4g (emulation vs fp32 %.4g)" % (
        k, ga.norm(), gb.norm(), (ga - gb).norm() / max(ga.norm().item(), 1e-20), (ge - gb).norm() / max(ge.norm().item(), 1e-20),
        (ge - ga).norm() / max(ga.norm().item(), 1e-20)))
