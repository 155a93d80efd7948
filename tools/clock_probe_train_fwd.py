"""Does the NeRF training forward (tile + mask checkpoints) run at lower SM clocks than the inference kernel?  Both looped for
~2 s on 1,048,576 rows while nvidia-smi samples clocks / power."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msra_practice_project_b200 import models, ops
torch.manual_seed(0)
net = models.NeRF().cuda()
n, s = 4096, 256
g = torch.Generator().manual_seed(1)
o = torch.tensor([0.0, 0.0, 4.0]).expand(n, 3)
d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
rays = torch.stack([o, d], 1).cuda()
z = (torch.sort(torch.rand(n, s, generator=g), -1).values * 4 + 2).cuda()
flat = models.flat_params(net).detach()
packed = ops.pack_tc(flat, models.KIND_NERF)

def sample(stop, out):
    while not stop[0]:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
        try:
            c, p = r.stdout.strip().split(","); out.append((float(c), float(p)))
        except Exception:
            pass
        time.sleep(0.1)

for name, fn in (("inference kernel", lambda: ops.mlp(net, rays=rays, z=z, precision="bf16")),
                 ("training forward", lambda: ops.tc_train_forward(packed, models.KIND_NERF, rays, z))):
    with torch.no_grad():
        for _ in range(20): fn()
        torch.cuda.synchronize()
        stop, out = [False], []
        th = threading.Thread(target=sample, args=(stop, out)); th.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 1500
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        stop[0] = True; th.join()
    clk = sorted(c for c, _ in out); pw = sorted(p for _, p in out)
    print("%s: %.3f ms per launch; SM clock median %.0f MHz, power median %.0f W (%d samples)" % (
        name, e0.elapsed_time(e1) / reps, clk[len(clk) // 2] if clk else -1, pw[len(pw) // 2] if pw else -1, len(out)))
