"""Kernel-level time breakdown of one UNet forward (plain vs fused) with torch.profiler (CUPTI)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from bndm_b200.fused_unet import fuse_unet
from bndm_b200.unet import get_model

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
which = sys.argv[2] if len(sys.argv) > 2 else "fused"        # plain | fused | fused_uniform (one timestep for the batch)
res = int(sys.argv[3]) if len(sys.argv) > 3 else 64          # 64 | 128 (cfg 4) | 512 = the latent UNet of cfg 5 (64^2 latents, 4 channels)
torch.manual_seed(0)
if res == 512:
    from bndm_b200.unet import get_latent_model
    model, cin, hw = get_latent_model(512, 8).to(dev).eval(), 4, 64
else:
    model, cin, hw = get_model(3, 6, res).to(dev).eval(), 3, res
kw = {}
if which.startswith("fused"):
    model = fuse_unet(model)
    if which == "fused_uniform":
        kw = {"uniform_timestep": True}
x = torch.randn(B, cin, hw, hw, device=dev)
t = torch.full((B,), 0.5, device=dev)
with torch.no_grad():
    for _ in range(3):
        model(x, t, return_dict=False, **kw)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            model(x, t, return_dict=False, **kw)
        torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(((e.device_time_total / 3.0, e.count // 3, e.key) for e in ev if e.device_time_total > 0), reverse=True)
total = sum(r[0] for r in rows)
print(f"{which} UNet forward, res={res} B={B}: {total / 1e3:.3f} ms of kernel time per forward, {sum(r[1] for r in rows)} launches")
for us, n, key in rows[:28]:
    print(f"{us:10.1f} us {100 * us / total:5.1f}%  x{n:<4d} {key[:110]}")
