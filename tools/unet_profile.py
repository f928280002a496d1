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
torch.manual_seed(0)
model = get_model(3, 6, 64).to(dev).eval()
kw = {}
if which.startswith("fused"):
    model = fuse_unet(model)
    if which == "fused_uniform":
        kw = {"uniform_timestep": True}
x = torch.randn(B, 3, 64, 64, device=dev)
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
print(f"{which} UNet forward, B={B}: {total / 1e3:.3f} ms of kernel time per forward, {sum(r[1] for r in rows)} launches")
for us, n, key in rows[:28]:
    print(f"{us:10.1f} us {100 * us / total:5.1f}%  x{n:<4d} {key[:110]}")
