"""Channels-last execution of the UNet with the normalisation glue fused into one kernel (K5).

The reference's sampler spends its time in ``model(x, t)`` (iadb_bn.py:319): a diffusers
``UNet2DModel`` -- here the plain-PyTorch restatement ``bndm_b200.unet.UNet2DModel``.  Run the
PyTorch way (NCHW fp32), one forward at batch 64 is ~700 launches of which the convolutions
proper are a quarter of the time; the rest is layout conversion around every cuDNN call
(nchw<->nhwc), GroupNorm as three kernels (moments, affine, SiLU) and separate bias / time-
embedding / residual adds (profiles/r01_launches_bench_1step.txt).

``FusedUNet2D`` wraps an existing ``UNet2DModel`` (sharing nothing mutable: it deep-copies the
weights into channels-last form) and evaluates the same network with
  * activations kept channels-last end to end (cuDNN's native layout: no conversions),
  * every ``SiLU(GroupNorm(.))`` -- and the ``+ time_emb_proj(temb)`` / conv1-bias adds in front
    of norm2 -- as ONE launch of ``bndm_groupnorm_nhwc_f32`` (csrc/groupnorm.cu),
  * conv1's bias folded into the per-sample time-embedding vector (a (B, C) add instead of a
    full activation pass).
Same call conventions as the wrapped model (``model(x, t, return_dict=False)[0]`` /
``.sample``); inference only (no autograd through K5).  Results agree with the wrapped module
to fp32 round-off of the normalisation (the convolutions are the same cuDNN TF32 kernels).
"""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F

from . import _lib
from .unet import UNet2DModel, UNet2DOutput, timestep_embedding


def groupnorm_silu_nhwc(x, norm, add_bc=None, res=None, want_sum=False, silu=True):
    """y = act(GroupNorm(x (+ res) (+ add_bc[:, :, None, None]))) on a channels-last (B,C,H,W) tensor.
    Returns y, or (y, s) with s = the pre-normalisation sum when ``want_sum``."""
    B, C, H, W = x.shape
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    if res is not None and not res.is_contiguous(memory_format=torch.channels_last):
        res = res.contiguous(memory_format=torch.channels_last)
    if x.dtype != torch.float32 or not x.is_cuda:
        raise _lib.BndmError("groupnorm_silu_nhwc: CUDA float32 tensors only (no CPU fallback)")
    y = torch.empty_like(x, memory_format=torch.channels_last)
    s = torch.empty_like(x, memory_format=torch.channels_last) if want_sum else None
    if add_bc is not None:
        add_bc = add_bc.contiguous()
    with torch.cuda.device(x.device):
        rc = _lib.load().bndm_groupnorm_nhwc_f32(_lib.ptr(x), _lib.ptr(res), _lib.ptr(add_bc), _lib.ptr(norm.weight),
                                                 _lib.ptr(norm.bias), _lib.ptr(s), _lib.ptr(y), B, C, H * W, norm.num_groups,
                                                 float(norm.eps), 1 if silu else 0, _lib.current_stream(x.device))
    _lib.check(rc, "bndm_groupnorm_nhwc_f32")
    return (y, s) if want_sum else y


class FusedUNet2D(torch.nn.Module):
    def __init__(self, model: UNet2DModel):
        super().__init__()
        if not isinstance(model, UNet2DModel):
            raise TypeError("FusedUNet2D wraps bndm_b200.unet.UNet2DModel")
        p = next(model.parameters())
        if p.dtype != torch.float32 or not p.is_cuda:
            raise _lib.BndmError("FusedUNet2D needs a float32 model on a CUDA device")
        self.m = copy.deepcopy(model).eval().to(memory_format=torch.channels_last)
        for q in self.m.parameters():
            q.requires_grad_(False)
        self.in_channels, self.out_channels = model.in_channels, model.out_channels

    # -- blocks ---------------------------------------------------------------------------------
    @staticmethod
    def _resnet(blk, x, temb_act):
        y = groupnorm_silu_nhwc(x, blk.norm1)
        h = F.conv2d(y, blk.conv1.weight, None, padding=1)
        tb = F.linear(temb_act, blk.time_emb_proj.weight, blk.time_emb_proj.bias) + blk.conv1.bias    # (B, Cout)
        y2 = groupnorm_silu_nhwc(h, blk.norm2, add_bc=tb)
        h2 = blk.conv2(y2)
        if blk.conv_shortcut is not None:
            x = blk.conv_shortcut(x)
        return x + h2

    def _down(self, block, h, temb_act):
        skips = []
        for i, resnet in enumerate(block.resnets):
            h = self._resnet(resnet, h, temb_act)
            if block.attentions is not None:
                h = block.attentions[i](h).contiguous(memory_format=torch.channels_last)
            skips.append(h)
        if block.downsamplers is not None:
            h = block.downsamplers[0](h)
            skips.append(h)
        return h, skips

    def _up(self, block, h, skips, temb_act):
        for i, resnet in enumerate(block.resnets):
            h = self._resnet(resnet, torch.cat([h, skips.pop()], dim=1), temb_act)
            if block.attentions is not None:
                h = block.attentions[i](h).contiguous(memory_format=torch.channels_last)
        if block.upsamplers is not None:
            h = block.upsamplers[0](h)
        return h

    @torch.no_grad()
    def forward(self, sample, timestep, return_dict=True):
        m = self.m
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.float32 if isinstance(t, float) else torch.int64, device=sample.device)
        elif t.dim() == 0:
            t = t[None].to(sample.device)
        t = t * torch.ones(sample.shape[0], dtype=t.dtype, device=t.device)
        emb = timestep_embedding(t, m.time_proj_dim)
        temb_act = F.silu(m.time_embedding(emb))

        h = m.conv_in(sample.float().contiguous(memory_format=torch.channels_last))
        skips = [h]
        for block in m.down_blocks:
            h, s = self._down(block, h, temb_act)
            skips.extend(s)
        h = self._resnet(m.mid_block.resnets[0], h, temb_act)
        h = m.mid_block.attentions[0](h).contiguous(memory_format=torch.channels_last)
        h = self._resnet(m.mid_block.resnets[1], h, temb_act)
        for block in m.up_blocks:
            h = self._up(block, h, skips, temb_act)
        h = m.conv_out(groupnorm_silu_nhwc(h, m.conv_norm_out))
        h = h.contiguous()                                   # NCHW for the step kernels
        if not return_dict:
            return (h,)
        return UNet2DOutput(sample=h)


def fuse_unet(model):
    """Returns the channels-last / fused-normalisation evaluator of ``model`` (a copy of its weights)."""
    return FusedUNet2D(model)
