"""One timed sampling run for BASELINE configs 3-5 on one GPU (per-GPU shard sizes), fused UNet + CUDA graph:
  cfg3  DDIM church_res64 UNet (out 3), 100 steps, eta = 0, B = 64
  cfg4  IADB cat_res128 UNet, 250 steps, B = 32 (= 256 over 8 GPUs), gamma sigmoid tau = 0.2, incl. get_noise_v2 at 128^2
  cfg5  latent IADB (4 x 64 x 64), 250 steps, B = 16 (= 128 over 8 GPUs), incl. get_noise_v2
Prints images/s per GPU (synthetic data, random-init weights)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bndm_b200 as bb
from bndm_b200.fused_unet import fuse_unet
from bndm_b200.synth import hashed_tril
from bndm_b200.unet import get_latent_model, get_model

dev = torch.device("cuda:0")
L = torch.from_numpy(hashed_tril(seed=0)).to(dev)
h = bb.prepare_L(L, max_columns=32 * 3 * 4)


def timed(fn, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / 1e3


torch.manual_seed(0)
with torch.no_grad():
    m3 = fuse_unet(get_model(3, 3, 64).to(dev).eval())
    x = torch.randn(64, 3, 64, 64, device=dev)
    s = timed(lambda: bb.sample_ddim(m3, x, 100, eta=0.0, use_graph=True))
    print(f"cfg3 DDIM res64 100 steps B=64: {s:.3f} s per run = {64 / s:.1f} images/s", flush=True)
    del m3

    m5 = fuse_unet(get_latent_model(512, 8).to(dev).eval())
    z = torch.randn(16, 4, 64, 64, device=dev)
    g1 = torch.ones(16, device=dev)

    def run5():
        x0 = bb.get_noise_v2(dev, z, h, g1, None, "gaussianBN", "test", True, want=("noise",))[0]
        return bb.sample_latent_iadb(m5, x0, 250, "gaussianBN", 8, use_graph=True)
    s = timed(run5)
    print(f"cfg5 latent IADB 4x64x64 250 steps B=16: {s:.3f} s per run = {16 / s:.1f} images/s", flush=True)
    del m5

    m4 = fuse_unet(get_model(3, 6, 128).to(dev).eval())
    w = torch.randn(32, 3, 128, 128, device=dev)
    g4 = torch.ones(32, device=dev)

    def run4():
        x0 = bb.get_noise_v2(dev, w, h, g4, None, "gaussianBN", "test", True, want=("noise",))[0]
        return bb.sample_iadb(m4, x0, 250, "sigmoid", (0.2, 0.0, 3.0), 6, "gaussianBN", "train", use_graph=True)
    s = timed(run4)
    print(f"cfg4 IADB res128 250 steps B=32: {s:.3f} s per run = {32 / s:.1f} images/s", flush=True)
