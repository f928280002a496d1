"""Artefact formats of the reference's drivers + image post-processing.

  cov_mat_L   np.load('./bluenoise/cov_gaussian{BN,RN}_L_res64_d3.npz')['x'].astype(float32)
              (iadb_bn.py:83-86, latent_iadb_bn_diffusers.py:63-67, gradio_bndm.py:55-58)
  x0 / noise  '<...>/noise/noise_batch{bs}_idx{:0>5}.npz'['noise']  (iadb_bn.py:764, ddim_diffusers.py:667)
  images      (x/2 + 0.5).clamp(0,1) -> NHWC -> *255 -> round -> uint8   (ddim_diffusers.py:687-688)
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def load_cov_mat_L(noise_type="gaussianBN", root="./bluenoise", device="cuda"):
    """Loads the reference's L file for this noise type; 'gaussianRN' selects the red-noise
    factor (iadb_bn.py:84-85), everything else the blue-noise one."""
    kind = "RN" if noise_type == "gaussianRN" else "BN"
    L = np.load(f"{root}/cov_gaussian{kind}_L_res64_d3.npz")["x"].astype(np.float32)
    if L.shape != (4096, 4096):
        raise ValueError(f"cov_mat_L must be (4096,4096), got {L.shape}")
    return torch.from_numpy(L).to(device).detach()


def noise_batch_path(root, batch_size, idx):
    return f"{root}/noise/noise_batch{batch_size}_idx{idx:0>5}.npz"


def load_noise_batch(root, batch_size, idx, device="cuda"):
    return torch.from_numpy(np.load(noise_batch_path(root, batch_size, idx))["noise"]).float().to(device)


def save_noise_batch(root, batch_size, idx, noise):
    np.savez(noise_batch_path(root, batch_size, idx), noise=noise.detach().cpu().numpy())


def to_uint8_nhwc(x: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) fp32 in [-1,1] -> (B,H,W,C) uint8, one fused kernel (K4)."""
    x = _lib.require_cuda_f32(x, "x")
    B, C, H, W = x.shape
    out = torch.empty((B, H, W, C), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().bndm_to_uint8_nhwc(_lib.ptr(x), _lib.ptr(out), B, C, H, W, _lib.current_stream(x.device))
    _lib.check(rc, "bndm_to_uint8_nhwc")
    return out


def iadb_snapshots_uint8(samples: torch.Tensor, final) -> torch.Tensor:
    """The same conversion for a stack of images ON THE DEVICE: (N,C,H,W) fp32 CUDA -> (N,H,W,C) uint8 CUDA, one kernel
    (one CTA per image: min/max reduction + conversion; replaces 4-6 torch kernels and the D2H copy of every fp32
    snapshot).  ``final``: bool for all images, or a length-N sequence / tensor (True = clamp((x+1)/2), False = min-max)."""
    x = _lib.require_cuda_f32(samples, "samples")
    if x.dim() != 4:
        raise ValueError("samples must be (N, C, H, W)")
    N, C, H, W = x.shape
    flags = None
    if not isinstance(final, bool):
        flags = torch.as_tensor(final).to(torch.int32).reshape(-1).to(x.device)
        if flags.numel() != N:
            raise ValueError(f"final must have {N} entries")
    out = torch.empty((N, H, W, C), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().bndm_snapshot_uint8_hwc(_lib.ptr(x), _lib.ptr(out), N, C, H, W, _lib.ptr(flags),
                                                 1 if (isinstance(final, bool) and final) else 0, _lib.current_stream(x.device))
    _lib.check(rc, "bndm_snapshot_uint8_hwc")
    return out


def iadb_snapshot_uint8(sample_chw: torch.Tensor, final: bool) -> np.ndarray:
    """(C,H,W) fp32 -> (H,W,C) uint8 exactly as iadb_bn.py's test driver writes its PNGs (:796-802, :814-816):
    the final image is ``clamp((x+1)/2, 0, 1)``, intermediate snapshots are min-max normalised, and the
    conversion is ``(x*255).astype(uint8)`` -- a TRUNCATION, unlike ddim_diffusers.py:687-688 which rounds
    (that variant is ``to_uint8_nhwc``).  Host-side convenience (plain torch ops, any device)."""
    x = sample_chw
    if final:
        x = torch.clamp((x + 1) / 2.0, 0.0, 1.0)
    else:
        x = (x - x.min()) / (x.max() - x.min())
    return (x.permute(1, 2, 0).detach().cpu().numpy() * 255).astype(np.uint8)
