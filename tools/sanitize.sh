#!/bin/bash
# compute-sanitizer passes over the hand-written kernels ("race detection / sanitizers"): memcheck, racecheck
# (shared-memory hazards), synccheck (barrier misuse) on a small driver that calls every contraction kernel, the step
# kernels and the UNet glue once.  Usage (GPU box): bash tools/sanitize.sh > profiles/rNN_sanitizer.txt
cat > /tmp/bndm_sanitize_driver.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import bndm_b200 as bb
from bndm_b200.synth import hashed_tril
from bndm_b200.fused_unet import fuse_unet
from bndm_b200.unet import get_latent_model
dev = torch.device("cuda:0")
L = torch.from_numpy(hashed_tril(seed=0)).to(dev)
for res, B, C, inplace in ((64, 4, 3, True), (32, 3, 4, True), (128, 1, 3, True), (64, 2, 4, False), (64, 8, 3, True), (64, 24, 3, True)):
    x = torch.randn(B, C, res, res, device=dev)
    g = torch.rand(B, device=dev)
    for gemm in ("auto", "tc"):
        bb.get_noise_v2(dev, x, L, g, None, "gaussianBN", "test", inplace, gemm=gemm)
torch.manual_seed(0)
m = fuse_unet(get_latent_model(256, 8).to(dev).eval())
z = torch.randn(2, 4, 32, 32, device=dev)
bb.sample_latent_iadb(m, z, 2, "gaussianBN", 8)
bb.sample_ddim(m, z, 2) if False else None
from bndm_b200.io import iadb_snapshots_uint8, to_uint8_nhwc
iadb_snapshots_uint8(torch.randn(3, 3, 64, 64, device=dev), [False, False, True]); to_uint8_nhwc(torch.randn(2, 3, 16, 16, device=dev))
# K9 / K10 (tcgen05 GEMMs of the fused UNet) and K5's warp / cluster variants at small and ragged shapes
from bndm_b200.fused_unet import linear_tc, shortcut_residual_nhwc, groupnorm_silu_nhwc
cl = torch.channels_last
linear_tc(torch.randn(300, 96, device=dev), torch.randn(516, 96, device=dev), torch.randn(516, device=dev))
linear_tc(torch.randn(256, 512, device=dev), torch.randn(1536, 512, device=dev))
for (B, C1, C2, N, H) in ((1, 96, 32, 132, 7), (2, 128, 128, 128, 16), (1, 128, 0, 256, 16)):
    xa = torch.randn(B, C1, H, H, device=dev).contiguous(memory_format=cl)
    xb = torch.randn(B, C2, H, H, device=dev).contiguous(memory_format=cl) if C2 else None
    shortcut_residual_nhwc(xa, xb, torch.randn(N, C1 + C2, 1, 1, device=dev), torch.randn(B, N, H, H, device=dev).contiguous(memory_format=cl),
                           torch.randn(N, device=dev))
for (B, C, H) in ((2, 512, 2), (2, 256, 4), (1, 128, 8), (2, 128, 64), (1, 128, 128), (2, 256, 16)):
    groupnorm_silu_nhwc(torch.randn(B, C, H, H, device=dev).contiguous(memory_format=cl), torch.nn.GroupNorm(32, C).to(dev),
                        add_bc=torch.randn(B, C, device=dev))
torch.cuda.synchronize()
print("driver done")
PY
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/bndm_sanitize_driver.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|driver done|Error|hazard|Invalid|=========     at" | head -20
done
