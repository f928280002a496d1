#!/usr/bin/env python
"""TFLOP/s of the cuDNN TF32 convolutions the fused UNet runs (channels-last, 3x3 pad 1 unless noted), B=64 at 64^2:
where the 55 % of the forward that is library code stands against the tensor-core peak.   python tools/conv_probe.py [B]"""
import os, sys, statistics
import torch
import torch.nn.functional as F

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
shapes = [(128, 128, 64, 3), (256, 128, 64, 3), (128, 128, 32, 3), (256, 128, 32, 3), (128, 256, 16, 3), (256, 256, 16, 3), (384, 256, 16, 3),
          (512, 256, 16, 3), (256, 256, 8, 3), (512, 256, 8, 3), (256, 512, 4, 3), (512, 512, 4, 3), (1024, 512, 4, 3), (512, 512, 2, 3),
          (1024, 512, 2, 3), (256, 128, 64, 1), (128, 128, 64, 1)]
for cin, cout, H, k in shapes:
    x = torch.randn(B, cin, H, H, device=dev).contiguous(memory_format=torch.channels_last)
    w = torch.randn(cout, cin, k, k, device=dev).contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        F.conv2d(x, w, None, padding=k // 2)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=s):
        for _ in range(10):
            y = F.conv2d(x, w, None, padding=k // 2)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 10)
    us = statistics.median(ts)
    flops = 2.0 * B * H * H * cin * cout * k * k
    nbytes = 4.0 * B * H * H * (cin + cout)
    print(f"conv {k}x{k} {cin:5d}->{cout:4d} @ {H:3d}^2: {us:8.2f} us  {flops / us / 1e6:7.1f} TFLOP/s  (activation bytes {nbytes / us / 1e3:6.0f} GB/s)", flush=True)
