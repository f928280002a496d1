#!/usr/bin/env python
"""Drift of the fused (channels-last, K5..K8) UNet evaluation against the stock module over a FULL cfg-2 trajectory:
64x64 IADB, 250 steps, gaussianBN / out_channel 6, same x0, same weights.  Reports the rms / max difference of x after
selected steps, with TF32 convolutions allowed (torch default, what the bench runs) and with TF32 off (fusion round-off
only).  Usage: python tools/unet_drift.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bndm_b200 as bb  # noqa: E402
from bndm_b200.fused_unet import fuse_unet  # noqa: E402
from bndm_b200.synth import blue_noise_L  # noqa: E402
from bndm_b200.unet import get_model  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device("cuda:0")
    T, params = 250, (1000.0, 0.0, 3.0)
    torch.manual_seed(0)
    model = get_model(3, 6, 64).to(dev).eval()
    L = torch.from_numpy(blue_noise_L()).to(dev)
    white = torch.randn(B, 3, 64, 64, device=dev)
    x0 = bb.get_noise_v2(dev, white, L, torch.ones(B, device=dev), None, "gaussianBN", "test", True)[0]
    marks = (249, 200, 150, 100, 50, 10, 0)
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        fused = fuse_unet(model)
        snaps = {}
        for name, m in (("stock", model), ("fused", fused)):
            s = bb.IadbSampler(m, x0.shape, T, "sigmoid", params, 6, "gaussianBN", device=dev, graph="step")
            keep = {}
            s.run(x0, on_step=lambda t, x: keep.__setitem__(t, x.clone()) if t in marks else None)
            snaps[name] = keep
            del s
        print(f"conv TF32 {'allowed (torch default)' if tf32 else 'off'}; batch {B}; x rms at t=0: {snaps['stock'][0].pow(2).mean().sqrt().item():.4f}")
        for t in marks:
            d = snaps["fused"][t] - snaps["stock"][t]
            print(f"   after step t={t:3d} ({T - t:3d} updates): rms diff {d.pow(2).mean().sqrt().item():.3e}   max |diff| {d.abs().max().item():.3e}")


if __name__ == "__main__":
    main()
