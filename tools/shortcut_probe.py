#!/usr/bin/env python
"""K10 (1x1 shortcut convolution + residual + biases in one kernel) against the three kernels it replaces (two cuDNN
GEMMs + K6) at the up blocks' shapes: us per call in a CUDA graph and GB/s of its algorithmic bytes.
    python tools/shortcut_probe.py [B]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from bndm_b200.fused_unet import shortcut_residual_nhwc, add_bias_residual_nhwc

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cl = torch.channels_last


def graph_us(fn, reps=10):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=s):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return statistics.median(ts)


for C1, C2, N, H in [(128, 128, 128, 64), (256, 128, 128, 32), (128, 128, 128, 32), (256, 256, 256, 16), (256, 128, 256, 16), (512, 256, 256, 8),
                     (512, 512, 512, 4), (512, 512, 512, 2)]:
    x = torch.randn(B, C1, H, H, device=dev).contiguous(memory_format=cl)
    x2 = torch.randn(B, C2, H, H, device=dev).contiguous(memory_format=cl)
    w = (torch.randn(N, C1 + C2, 1, 1, device=dev) / (C1 + C2) ** 0.5).contiguous(memory_format=cl)
    w1, w2 = w[:, :C1].contiguous(memory_format=cl), w[:, C1:].contiguous(memory_format=cl)
    h2 = torch.randn(B, N, H, H, device=dev).contiguous(memory_format=cl)
    b1, b2 = torch.randn(N, device=dev), torch.randn(N, device=dev)
    bt = b1 + b2
    t_new = graph_us(lambda: shortcut_residual_nhwc(x, x2, w, h2, bt))
    t_old = graph_us(lambda: add_bias_residual_nhwc(F.conv2d(x, w1), h2, b2, bias_a=b1, a2=F.conv2d(x2, w2)))
    nbytes = 4.0 * B * H * H * (C1 + C2 + 2 * N)
    print(f"{C1:4d}+{C2:4d} -> {N:4d} @ {H:3d}^2: K10 {t_new:8.2f} us ({nbytes / t_new / 1e3:6.0f} GB/s)   two cuDNN GEMMs + K6 {t_old:8.2f} us", flush=True)
