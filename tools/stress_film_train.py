"""Stress of the FiLM-SIREN fused training path: random row counts / latent counts, forward + backward, finiteness and
agreement of the batched call with per-latent calls.  Guards against barrier-protocol hangs at odd tile counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msra_practice_project_b200 import models, ops
torch.manual_seed(0)
net = models.FilmSirenNeRF().cuda()
g = torch.Generator().manual_seed(3)
t0 = time.time()
for it in range(40):
    b = int(torch.randint(1, 8, (1,), generator=g))
    rpl = 512 * int(torch.randint(1, 60, (1,), generator=g))
    s = [1, 2, 4, 8, 16][int(torch.randint(0, 5, (1,), generator=g))]
    n = b * rpl // s
    film = torch.cat([1.0 + 0.1 * torch.randn(b, 9, 256, generator=g), 0.1 * torch.randn(b, 9, 256, generator=g)], -1).cuda().requires_grad_(True)
    o = torch.tensor([0.0, 0.0, 1.0]).expand(n, 3)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.25 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
    rays = torch.stack([o, d], 1).cuda()
    z = (torch.sort(torch.rand(n, s, generator=g), -1).values + 0.5).cuda()
    up = torch.randn(n * s, 4, generator=g).cuda()
    net.zero_grad(set_to_none=True)
    raw = ops.mlp_film_batched_train(net, film, rays, z, rpl)
    (raw * up).sum().backward()
    torch.cuda.synchronize()
    gw = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.isfinite(raw).all() and torch.isfinite(gw).all() and torch.isfinite(film.grad).all()
    # latent 0 alone must reproduce its rows and its d film
    f0 = film[0].detach().clone().requires_grad_(True)
    net.zero_grad(set_to_none=True)
    net.set_film_params(f0)
    rows0 = rpl // s
    raw0 = ops.mlp(net, rays=rays[:rows0], z=z[:rows0])
    (raw0 * up[:rpl]).sum().backward()
    torch.cuda.synchronize()
    assert torch.equal(raw0, raw[:rpl].detach()), it
    rel = (f0.grad - film.grad[0]).norm().item() / max(f0.grad.norm().item(), 1e-20)
    assert rel < 1e-3, (it, rel)
    print(f"iter {it}: B={b} rows/latent={rpl} S={s} ok ({time.time() - t0:.1f} s)", flush=True)
print("stress ok")
