#!/usr/bin/env python
"""Does cuDNN's autotuner (torch.backends.cudnn.benchmark) pick faster convolution kernels than its heuristics for the
fused UNet?  One forward in a CUDA graph, ms per forward, benchmark off / on.   python tools/cudnn_benchmark_probe.py [B] [res]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bndm_b200.fused_unet import fuse_unet
from bndm_b200.unet import get_model

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
res = int(sys.argv[2]) if len(sys.argv) > 2 else 64
for bench in (False, True, False):
    torch.backends.cudnn.benchmark = bench
    torch.manual_seed(0)
    model = fuse_unet(get_model(3, 6, res).to(dev).eval())
    x = torch.randn(B, 3, res, res, device=dev)
    t = torch.full((B,), 0.5, device=dev)
    with torch.no_grad():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                y = model(x, t, return_dict=False, uniform_timestep=True)[0]
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            y = model(x, t, return_dict=False, uniform_timestep=True)[0]
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) / 5)
    print(f"res={res} B={B} cudnn.benchmark={bench}: {statistics.median(ts):.3f} ms per forward  (checksum {y.double().sum().item():.6f})", flush=True)
