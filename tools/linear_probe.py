#!/usr/bin/env python
"""K9 (3xTF32 tcgen05 linear) against torch's fp32 F.linear at the attention blocks' shapes: us per call in a CUDA graph.
    python tools/linear_probe.py [B]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from bndm_b200.fused_unet import linear_tc

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64


def graph_us(fn, reps=20):
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


for M, N, K in [(B * 16, 1536, 512), (B * 16, 512, 512), (B * 64, 768, 256), (B * 4, 1536, 512)]:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    t_tc = graph_us(lambda: linear_tc(a, w, b))
    t_ref = graph_us(lambda: F.linear(a, w, b))
    err = (linear_tc(a, w, b).double() - F.linear(a.double(), w.double(), b.double())).abs().max().item()
    err32 = (F.linear(a, w, b).double() - F.linear(a.double(), w.double(), b.double())).abs().max().item()
    print(f"M={M:5d} N={N:5d} K={K:4d}: K9 {t_tc:7.2f} us ({2.0 * M * N * K / t_tc / 1e6:6.1f} TFLOP/s)   torch fp32 {t_ref:7.2f} us   "
          f"max|err| vs fp64: K9 {err:.2e}, torch fp32 {err32:.2e}", flush=True)
