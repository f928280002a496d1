"""Plain-PyTorch ``UNet2DModel`` with the architecture and state-dict key names of the
diffusers model the reference constructs (iadb_bn.py:205-282, utils.py:7-84,
ddim_diffusers.py:375-458, latent_iadb_bn_diffusers.py:334-372).

diffusers is not installed offline, so the network the samplers drive is restated here
from the published architecture (parity-unpinned against diffusers
source, but module/parameter names follow diffusers so a real ``model.ckpt``
(iadb_bn.py:714) or ``unet/*.safetensors`` state dict loads with ``load_state_dict``).
The UNet forward stays in PyTorch (cuDNN / cuBLAS); this repo wraps it in CUDA graphs and
feeds / consumes it with its own kernels.

Call conventions kept: ``model(x, t, return_dict=False)[0]`` (iadb_bn.py:319) and
``model(x, t).sample`` (ddim_diffusers.py:679); ``t`` may be a python number, a 0-dim
tensor or a (B,) tensor, int or float (IADB feeds alpha in (0,1]).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNet2DOutput:
    sample: torch.Tensor


def timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos=True, freq_shift=0.0, max_period=10000):
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb_act):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(temb_act)[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """Spatial self-attention block (diffusers ``Attention`` in its AttnBlock configuration):
    GroupNorm -> q,k,v Linear -> softmax(QK^T/sqrt(d))V with d = head dim 8 -> Linear, residual."""

    def __init__(self, channels, head_dim=8, groups=32, eps=1e-5):
        super().__init__()
        self.heads = channels // head_dim
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x.reshape(B, C, H * W)).transpose(1, 2)            # (B, HW, C)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)

        def split(t):
            return t.reshape(B, H * W, self.heads, C // self.heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(split(q), split(k), split(v))
        o = o.transpose(1, 2).reshape(B, H * W, C)
        o = self.to_out[0](o)
        return x + o.transpose(1, 2).reshape(B, C, H, W)


class Downsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers, add_downsample, attention, head_dim):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels)
                                      for i in range(num_layers)])
        self.attentions = nn.ModuleList([Attention(out_channels, head_dim) for _ in range(num_layers)]) if attention else None
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, h, temb_act):
        skips = []
        for i, resnet in enumerate(self.resnets):
            h = resnet(h, temb_act)
            if self.attentions is not None:
                h = self.attentions[i](h)
            skips.append(h)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            skips.append(h)
        return h, skips


class UpBlock2D(nn.Module):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers, add_upsample,
                 attention, head_dim):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            skip_ch = in_channels if i == num_layers - 1 else out_channels
            res_in = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(res_in + skip_ch, out_channels, temb_channels))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = nn.ModuleList([Attention(out_channels, head_dim) for _ in range(num_layers)]) if attention else None
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, h, skips, temb_act):
        for i, resnet in enumerate(self.resnets):
            h = resnet(torch.cat([h, skips.pop()], dim=1), temb_act)
            if self.attentions is not None:
                h = self.attentions[i](h)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels, temb_channels, head_dim):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels, temb_channels) for _ in range(2)])
        self.attentions = nn.ModuleList([Attention(channels, head_dim)])

    def forward(self, h, temb_act):
        h = self.resnets[0](h, temb_act)
        h = self.attentions[0](h)
        return self.resnets[1](h, temb_act)


class UNet2DModel(nn.Module):
    def __init__(self, sample_size=None, in_channels=3, out_channels=3,
                 down_block_types=("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
                 up_block_types=("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
                 block_out_channels=(224, 448, 672, 896), layers_per_block=2, act_fn="silu", attention_head_dim=8,
                 norm_num_groups=32, add_attention=True):
        super().__init__()
        if norm_num_groups != 32:
            # every block of this restatement is built with the diffusers default of 32 groups (the only value the
            # reference's configs use, iadb_bn.py:205-282): refuse instead of silently ignoring another value
            raise NotImplementedError("UNet2DModel: norm_num_groups other than 32 is not implemented")
        if act_fn != "silu":
            raise NotImplementedError("the reference only uses act_fn='silu' (iadb_bn.py:60,282)")
        if len(down_block_types) != len(up_block_types) or len(down_block_types) != len(block_out_channels):
            raise ValueError("block type / channel tuples must have equal length")
        self.sample_size = sample_size
        self.in_channels, self.out_channels = in_channels, out_channels
        ch0 = block_out_channels[0]
        temb = ch0 * 4
        self.time_proj_dim = ch0
        self.time_embedding = TimestepEmbedding(ch0, temb)
        self.conv_in = nn.Conv2d(in_channels, ch0, 3, padding=1)

        self.down_blocks = nn.ModuleList()
        out_ch = ch0
        for i, kind in enumerate(down_block_types):
            in_ch, out_ch = out_ch, block_out_channels[i]
            self.down_blocks.append(DownBlock2D(in_ch, out_ch, temb, layers_per_block,
                                                add_downsample=i != len(block_out_channels) - 1,
                                                attention=kind == "AttnDownBlock2D", head_dim=attention_head_dim))
        self.mid_block = UNetMidBlock2D(block_out_channels[-1], temb, attention_head_dim)

        self.up_blocks = nn.ModuleList()
        rev = list(reversed(block_out_channels))
        out_ch = rev[0]
        for i, kind in enumerate(up_block_types):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(rev) - 1)]
            self.up_blocks.append(UpBlock2D(in_ch, prev, out_ch, temb, layers_per_block + 1,
                                            add_upsample=i != len(rev) - 1,
                                            attention=kind == "AttnUpBlock2D", head_dim=attention_head_dim))
        self.conv_norm_out = nn.GroupNorm(min(ch0 // 4, 32), ch0, eps=1e-5)
        self.conv_out = nn.Conv2d(ch0, out_channels, 3, padding=1)

    def forward(self, sample, timestep, return_dict=True):
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.float32 if isinstance(t, float) else torch.int64, device=sample.device)
        elif t.dim() == 0:
            t = t[None].to(sample.device)
        t = t * torch.ones(sample.shape[0], dtype=t.dtype, device=t.device)
        emb = timestep_embedding(t, self.time_proj_dim).to(self.conv_in.weight.dtype)
        temb_act = F.silu(self.time_embedding(emb))       # every resnet applies SiLU to temb first

        h = self.conv_in(sample.to(self.conv_in.weight.dtype))
        skips = [h]
        for block in self.down_blocks:
            h, s = block(h, temb_act)
            skips.extend(s)
        h = self.mid_block(h, temb_act)
        for block in self.up_blocks:
            h = block(h, skips, temb_act)
        h = self.conv_out(F.silu(self.conv_norm_out(h)))
        h = h.float().contiguous()
        if not return_dict:
            return (h,)
        return UNet2DOutput(sample=h)


_CONFIGS = {
    # res -> block_out_channels; attention sits in the second-to-last down block / second up block
    64: (128, 128, 256, 256, 512, 512),
    128: (128, 128, 128, 256, 256, 512, 512),
    256: (128, 128, 128, 128, 256, 256, 512, 512),
}


def get_model(inp_channel=3, out_channel=3, res=64):
    """utils.get_model / iadb_bn.get_model (utils.py:7-84): the 64 / 128 / 256 pixel UNets."""
    if res not in _CONFIGS:
        raise NotImplementedError
    chans = _CONFIGS[res]
    n = len(chans)
    down = tuple("AttnDownBlock2D" if i == n - 2 else "DownBlock2D" for i in range(n))
    up = tuple("AttnUpBlock2D" if i == 1 else "UpBlock2D" for i in range(n))
    return UNet2DModel(block_out_channels=chans, out_channels=out_channel, in_channels=inp_channel,
                       up_block_types=up, down_block_types=down, act_fn="silu", add_attention=True)


def get_latent_model(resolution=512, out_channels=8):
    """latent_iadb_bn_diffusers.py:334-372 (in_channels=4; out_channels already doubled for BN/RN, :282)."""
    table = {
        64: ((128, 128, 256, 256, 512, 512), 4, 1), 512: ((128, 128, 256, 256, 512, 512), 4, 1),
        128: ((128, 128, 128, 256, 256, 512, 512), 5, 1),
        256: ((128, 256, 256), 2, 0),
    }
    if resolution not in table:
        raise ValueError(f"Unsupported resolution: {resolution}")
    chans, attn_down, attn_up = table[resolution]
    n = len(chans)
    down = tuple("AttnDownBlock2D" if i == attn_down else "DownBlock2D" for i in range(n))
    up = tuple("AttnUpBlock2D" if i == attn_up else "UpBlock2D" for i in range(n))
    return UNet2DModel(sample_size=resolution, in_channels=4, out_channels=out_channels, layers_per_block=2,
                       block_out_channels=chans, down_block_types=down, up_block_types=up)


def count_forward_flops(model: UNet2DModel, H: int, W: int) -> float:
    """Analytic conv + linear + attention FLOPs (MAC x 2) of one forward for one image."""
    total = 0.0
    hooks = []

    def conv_hook(m, inp, out):
        nonlocal total
        k = m.kernel_size[0] * m.kernel_size[1]
        total += 2.0 * out.shape[1] * out.shape[2] * out.shape[3] * m.in_channels * k / m.groups

    def lin_hook(m, inp, out):
        nonlocal total
        total += 2.0 * (out.numel() // out.shape[0]) * m.in_features

    def attn_hook(m, inp, out):
        nonlocal total
        _, C, h, w = inp[0].shape
        total += 4.0 * (h * w) ** 2 * C

    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            hooks.append(m.register_forward_hook(conv_hook))
        elif isinstance(m, nn.Linear):
            hooks.append(m.register_forward_hook(lin_hook))
        elif isinstance(m, Attention):
            hooks.append(m.register_forward_hook(attn_hook))
    p = next(model.parameters())
    with torch.no_grad():
        model(torch.zeros(1, model.in_channels, H, W, device=p.device, dtype=p.dtype), torch.ones(1, device=p.device))
    for h_ in hooks:
        h_.remove()
    return total
