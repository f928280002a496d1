"""K5 (fused GroupNorm+SiLU, NHWC) at the UNet's shapes, B=64: us per call and achieved GB/s
(algorithmic bytes = read x once + write y once)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bndm_b200.fused_unet import groupnorm_silu_nhwc

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
shapes = [(128, 64), (256, 64), (384, 64), (128, 32), (256, 32), (384, 32), (256, 16), (512, 16), (768, 16), (256, 8), (512, 8),
          (512, 4), (768, 4), (1024, 4), (512, 2), (1024, 2)]
if os.environ.get("K5_SHAPES"):
    shapes = [tuple(int(v) for v in t.split("x")) for t in os.environ["K5_SHAPES"].split(",")]
for C, H in shapes:
    xs = [torch.randn(B, C, H, H, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(4)]
    norm = torch.nn.GroupNorm(32, C).to(dev)
    tb = torch.randn(B, C, device=dev)
    for _ in range(2):
        groupnorm_silu_nhwc(xs[0], norm, add_bc=tb)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=s):
        for i in range(20):
            groupnorm_silu_nhwc(xs[i % 4], norm, add_bc=tb)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 20)
    us = statistics.median(ts)
    nbytes = 2 * B * C * H * H * 4
    tt = [float("nan")]
    for _ in range(0 if os.environ.get("K5_NO_TORCH") else 5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            torch.nn.functional.silu(norm(xs[i % 4]))
        e1.record(); e1.synchronize()
        tt.append(e0.elapsed_time(e1) * 1e3 / 20)
    print(f"C={C:5d} H={H:3d}: K5 {us:8.2f} us  {nbytes / us / 1e3:7.0f} GB/s   (torch channels-last GroupNorm+SiLU: {statistics.median(tt):8.2f} us)", flush=True)
