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
    full activation pass), all 30 ``time_emb_proj`` linears evaluated as ONE GEMM per forward,
  * conv2's bias and the residual add as one pass (``bndm_add_bias_nhwc_f32``, K6) -- or, where the shortcut is a 1x1
    convolution (every resnet of the up blocks), the shortcut GEMM of [h | skip], the residual add and both biases as ONE
    TF32 tcgen05 kernel (``bndm_shortcut_residual_tf32``, K10),
  * the attention blocks' fp32 q/k/v and output projections as a 3xTF32 tcgen05 GEMM (``bndm_linear_tc_f32``, K9),
  * ``conv_in`` from the sampler's NCHW state straight to the channels-last activation (``bndm_conv_in3x3_nhwc_f32``, K11).
K9 / K10 replace TF32-class library kernels and are used only while ``torch.backends.cudnn.allow_tf32`` is set (torch's
default and the reference's configuration); with TF32 switched off every GEMM stays torch's.
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


LAUNCHES = 0      # launches of libbndm_b200.so kernels issued from this module (host-side count)
TIMING = None     # when a list: every K5 / K6 / K7 launch is bracketed by CUDA events on its stream and
                  # (name, algorithmic bytes, start event, end event, launch closure) is appended (bench.py's live
                  # roofline replays the closures -- same tensors -- from one CUDA graph)


def _timed_launch(name, nbytes, device, launch):
    if TIMING is None:
        return launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream(device))
    rc = launch()
    e1.record(torch.cuda.current_stream(device))
    TIMING.append((name, nbytes, e0, e1, launch))
    return rc


def groupnorm_silu_nhwc(x, norm, add_bc=None, res=None, want_sum=False, silu=True, x2=None):
    """y = act(GroupNorm(x (+ res) (+ add_bc[:, :, None, None]))) on a channels-last (B,C,H,W) tensor.
    Returns y, or (y, s) with s = the pre-normalisation sum when ``want_sum``.
    ``x2``: the input is ``torch.cat((x, x2), 1)`` -- read from the two tensors, never materialised."""
    B, C1, H, W = x.shape
    C = C1 + (x2.shape[1] if x2 is not None else 0)
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    if x2 is not None and not x2.is_contiguous(memory_format=torch.channels_last):
        x2 = x2.contiguous(memory_format=torch.channels_last)
    if res is not None and not res.is_contiguous(memory_format=torch.channels_last):
        res = res.contiguous(memory_format=torch.channels_last)
    if x.dtype != torch.float32 or not x.is_cuda:
        raise _lib.BndmError("groupnorm_silu_nhwc: CUDA float32 tensors only (no CPU fallback)")
    y = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    s = torch.empty_like(y) if want_sum else None
    stride = 0
    if add_bc is not None:
        if add_bc.dim() == 1:                                                          # per-channel bias: one row for all samples
            add_bc = add_bc[None, :].expand(B, C)
        if add_bc.dim() != 2 or add_bc.shape != (B, C) or add_bc.stride(1) != 1:      # a column slice is fine
            add_bc = add_bc.reshape(B, C).contiguous()
        stride = add_bc.stride(0)
    with torch.cuda.device(x.device):
        rc = _timed_launch("K5", 8 * y.numel(), x.device, lambda: _lib.load().bndm_groupnorm_nhwc_f32(
            _lib.ptr(x), _lib.ptr(x2), C1, _lib.ptr(res), _lib.ptr(add_bc), stride, _lib.ptr(norm.weight), _lib.ptr(norm.bias),
            _lib.ptr(s), _lib.ptr(y), B, C, H * W, norm.num_groups, float(norm.eps), 1 if silu else 0,
            _lib.current_stream(x.device)))
    if rc == _lib.ERR_UNSUPPORTED and x2 is not None:          # slab too large for the two-source kernel
        return groupnorm_silu_nhwc(torch.cat([x, x2], 1), norm, add_bc, res, want_sum, silu)
    _lib.check(rc, "bndm_groupnorm_nhwc_f32")
    global LAUNCHES
    LAUNCHES += 1
    return (y, s) if want_sum else y


def add_bias_residual_nhwc(a, b, bias, bias_a=None, a2=None):
    """((a [+ a2]) [+ bias_a]) + (b + bias) with per-channel biases, channels-last (B,C,H,W) tensors: conv
    biases + residual add in one pass (K6)."""
    if a2 is not None and not a2.is_contiguous(memory_format=torch.channels_last):
        a2 = a2.contiguous(memory_format=torch.channels_last)
    if not (a.is_contiguous(memory_format=torch.channels_last) and b.is_contiguous(memory_format=torch.channels_last)):
        a = a.contiguous(memory_format=torch.channels_last)
        b = b.contiguous(memory_format=torch.channels_last)
    out = torch.empty_like(a, memory_format=torch.channels_last)
    with torch.cuda.device(a.device):
        rc = _timed_launch("K6", 4 * a.numel() * (3 + (a2 is not None)), a.device, lambda: _lib.load().bndm_add_bias_nhwc_f32(
            _lib.ptr(a), _lib.ptr(a2), _lib.ptr(bias_a), _lib.ptr(b), _lib.ptr(bias), _lib.ptr(out), a.numel(), a.shape[1],
            _lib.current_stream(a.device)))
    _lib.check(rc, "bndm_add_bias_nhwc_f32")
    global LAUNCHES
    LAUNCHES += 1
    return out


def upsample2x_nhwc(x):
    """F.interpolate(x, scale_factor=2, mode="nearest") for a channels-last (B,C,H,W) tensor (K8)."""
    B, C, H, W = x.shape
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    y = torch.empty((B, C, 2 * H, 2 * W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        rc = _timed_launch("K8", 4 * (x.numel() + y.numel()), x.device, lambda: _lib.load().bndm_upsample2x_nhwc_f32(
            _lib.ptr(x), _lib.ptr(y), B, H, W, C, _lib.current_stream(x.device)))
    _lib.check(rc, "bndm_upsample2x_nhwc_f32")
    global LAUNCHES
    LAUNCHES += 1
    return y


def attention_small(qkv, C, head_dim=8):
    """softmax(q k^T / sqrt(d)) v for (B, T, 3C) packed projections with tiny T (K7)."""
    B, T, _ = qkv.shape
    qkv = qkv.contiguous()
    out = torch.empty(B, T, C, dtype=torch.float32, device=qkv.device)
    with torch.cuda.device(qkv.device):
        rc = _timed_launch("K7", 4 * (qkv.numel() + out.numel()), qkv.device, lambda: _lib.load().bndm_attention_small_f32(
            _lib.ptr(qkv), _lib.ptr(out), B, T, C, head_dim, _lib.current_stream(qkv.device)))
    _lib.check(rc, "bndm_attention_small_f32")
    global LAUNCHES
    LAUNCHES += 1
    return out


def linear_tc(x, w, bias=None):
    """F.linear(x, w, bias) for CUDA fp32 tensors on the tensor cores with fp32-grade results (K9, 3xTF32)."""
    K = x.shape[-1]
    a = x.reshape(-1, K)
    if not a.is_contiguous():
        a = a.contiguous()
    if not w.is_contiguous():
        w = w.contiguous()
    M, N = a.shape[0], w.shape[0]
    out = torch.empty(M, N, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _timed_launch("K9", 4 * (a.numel() + w.numel() + out.numel()), x.device, lambda: _lib.load().bndm_linear_tc_f32(
            _lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), M, N, K, _lib.current_stream(x.device)))
    _lib.check(rc, "bndm_linear_tc_f32")
    global LAUNCHES
    LAUNCHES += 1
    return out.reshape(*x.shape[:-1], N)


def shortcut_residual_nhwc(x, x2, w, h2, bias):
    """conv1x1(cat(x, x2), w) + h2 + bias on channels-last (B,C,H,W) tensors in one kernel (K10): the shortcut convolution
    of a ResnetBlock2D in TF32 on the tensor cores with the residual and both biases added in its epilogue."""
    B, C1, H, W = x.shape
    C2 = x2.shape[1] if x2 is not None else 0
    N = w.shape[0]
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    if x2 is not None and not x2.is_contiguous(memory_format=torch.channels_last):
        x2 = x2.contiguous(memory_format=torch.channels_last)
    if not h2.is_contiguous(memory_format=torch.channels_last):
        h2 = h2.contiguous(memory_format=torch.channels_last)
    w2 = w.reshape(N, C1 + C2)
    if not w2.is_contiguous():
        w2 = w2.contiguous()
    out = torch.empty_like(h2, memory_format=torch.channels_last)
    M = B * H * W
    with torch.cuda.device(x.device):
        rc = _timed_launch("K10", 4 * (M * (C1 + C2) + 2 * M * N), x.device, lambda: _lib.load().bndm_shortcut_residual_tf32(
            _lib.ptr(x), _lib.ptr(x2), C1, C2, _lib.ptr(w2), _lib.ptr(h2), _lib.ptr(bias), _lib.ptr(out), M, N,
            _lib.current_stream(x.device)))
    _lib.check(rc, "bndm_shortcut_residual_tf32")
    global LAUNCHES
    LAUNCHES += 1
    return out


def conv_in3x3_nhwc(x, w):
    """conv2d(x, w, padding=1) for an NCHW fp32 input with <= 4 channels, written as a channels-last (B,Cout,H,W) tensor (K11);
    ``w`` is the contiguous (Cout, Cin, 3, 3) weight.  None when the shape is not one the kernel takes."""
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    if Cin > 4 or Cout % 4 != 0 or 256 % (Cout // 4) != 0 or tuple(w.shape[1:]) != (Cin, 3, 3):
        return None
    x = x.contiguous()
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        rc = _timed_launch("K11", 4 * (x.numel() + out.numel()), x.device, lambda: _lib.load().bndm_conv_in3x3_nhwc_f32(
            _lib.ptr(x), _lib.ptr(w), _lib.ptr(out), B, Cin, H, W, Cout, _lib.current_stream(x.device)))
    _lib.check(rc, "bndm_conv_in3x3_nhwc_f32")
    global LAUNCHES
    LAUNCHES += 1
    return out


def _use_shortcut_tc(x, x2, w):
    """K10 replaces cuDNN's TF32 1x1 convolution, so only while TF32 convolutions are allowed (the reference's configuration),
    and where it is faster than the three kernels it replaces: from 8192 rows up (tools/shortcut_probe.py: 108.8 vs 171.7 us
    at 64^2 x 128 channels and batch 64, 19.6 vs 27.3 us at 16^2; below, a CTA's K loop is a latency chain and cuDNN wins)."""
    return (x.is_cuda and x.dtype == torch.float32 and torch.backends.cudnn.allow_tf32 and x.shape[1] % 32 == 0
            and (x2 is None or x2.shape[1] % 32 == 0) and w.shape[0] % 4 == 0 and x.shape[0] * x.shape[2] * x.shape[3] >= 8192)


def _use_linear_tc(x, w):
    """K9 where it pays (>= 256 rows) and where the convolutions around it run in TF32 anyway (the reference's
    configuration); with TF32 switched off the linears stay torch's fp32 GEMMs so the isolation tests compare like with like."""
    return (x.is_cuda and x.dtype == torch.float32 and torch.backends.cudnn.allow_tf32 and x.numel() // x.shape[-1] >= 256
            and x.shape[-1] % 32 == 0 and w.shape[0] % 4 == 0)


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
        # every resnet's time_emb_proj as ONE GEMM per forward: rows of W_all are the concatenated
        # projection weights; conv1's bias rides along (it is added at the same place, before norm2)
        blocks = self._resnets()
        self.temb_w = torch.cat([r.time_emb_proj.weight for r in blocks], 0).contiguous()
        self.temb_b = torch.cat([r.time_emb_proj.bias + r.conv1.bias for r in blocks], 0).contiguous()
        self._temb_slices, off = {}, 0
        for r in blocks:
            n = r.time_emb_proj.out_features
            self._temb_slices[id(r)] = (off, off + n)
            off += n

        self._conv_in_w = self.m.conv_in.weight.detach().contiguous().clone()      # (Cout, Cin, 3, 3) row-major for K11
        self._sc_w = {}
        self._pend = {}
        self._qkv = {}
        for mod in self.m.modules():
            if mod.__class__.__name__ == "Attention":
                self._qkv[id(mod)] = (torch.cat([mod.to_q.weight, mod.to_k.weight, mod.to_v.weight], 0).contiguous(),
                                      torch.cat([mod.to_q.bias, mod.to_k.bias, mod.to_v.bias], 0).contiguous())

    def _resnets(self):
        m = self.m
        out = []
        for blk in m.down_blocks:
            out.extend(blk.resnets)
        out.extend(m.mid_block.resnets)
        for blk in m.up_blocks:
            out.extend(blk.resnets)
        return out

    # -- blocks ---------------------------------------------------------------------------------
    def _resnet(self, blk, x, tb_all, x2=None, x_bias=None, x2_bias=None):
        """x2: the block's input is cat((x, x2), 1) (up blocks) -- never materialised: norm1 reads both
        tensors and the 1x1 conv_shortcut is applied to the two halves separately.
        x_bias / x2_bias: per-channel biases still owed to x / x2 (the producing convolution ran without
        its bias): norm1 adds them on the fly and the shortcut absorbs them into its own bias."""
        lo, hi = self._temb_slices[id(blk)]
        pend = self._pending(blk, x.shape[1], x_bias, x2.shape[1] if x2 is not None else 0, x2_bias)
        y = groupnorm_silu_nhwc(x, blk.norm1, x2=x2, add_bc=pend["norm_add"])
        h = F.conv2d(y, blk.conv1.weight, None, padding=1)                 # bias folded into tb_all
        y2 = groupnorm_silu_nhwc(h, blk.norm2, add_bc=tb_all[0, lo:hi] if tb_all.shape[0] == 1 else tb_all[:, lo:hi])
        h2 = F.conv2d(y2, blk.conv2.weight, None, padding=1)               # bias added with the residual (K6)
        if blk.conv_shortcut is not None and _use_shortcut_tc(x, x2, blk.conv_shortcut.weight):
            # K10: the 1x1 shortcut convolution of [x | x2], the residual add and both biases in one kernel
            if "tail_bias" not in pend:
                pend["tail_bias"] = (pend["sc_bias"] + blk.conv2.bias).contiguous()
            return shortcut_residual_nhwc(x, x2, blk.conv_shortcut.weight, h2, pend["tail_bias"])
        if x2 is not None:
            w1, w2 = self._sc_split(blk, x.shape[1])
            sc, sc2 = F.conv2d(x, w1, None), F.conv2d(x2, w2, None)
            return add_bias_residual_nhwc(sc, h2, blk.conv2.bias, bias_a=pend["sc_bias"], a2=sc2)
        if blk.conv_shortcut is not None:
            sc = F.conv2d(x, blk.conv_shortcut.weight, None)              # 1x1; its bias is added in K6 too
            return add_bias_residual_nhwc(sc, h2, blk.conv2.bias, bias_a=pend["sc_bias"])
        return add_bias_residual_nhwc(x, h2, blk.conv2.bias, bias_a=x_bias)

    def _pending(self, blk, c1, x_bias, c2, x2_bias):
        """Cached per block: the (C,) vector norm1 must add and the shortcut's effective bias."""
        key = id(blk)
        if key not in self._pend:
            norm_add, sc_bias = None, (blk.conv_shortcut.bias if blk.conv_shortcut is not None else None)
            if x_bias is not None or x2_bias is not None:
                dev = blk.norm1.weight.device
                parts = [x_bias if x_bias is not None else torch.zeros(c1, device=dev)]
                if c2:
                    parts.append(x2_bias if x2_bias is not None else torch.zeros(c2, device=dev))
                norm_add = torch.cat(parts).contiguous()
                if blk.conv_shortcut is not None:                         # W (x + b) = W x + W b
                    sc_bias = blk.conv_shortcut.bias + blk.conv_shortcut.weight[:, :, 0, 0] @ norm_add
            self._pend[key] = {"norm_add": norm_add, "sc_bias": sc_bias}
        return self._pend[key]

    def _sc_split(self, blk, c1):
        key = (id(blk), c1)
        if key not in self._sc_w:
            w = blk.conv_shortcut.weight
            self._sc_w[key] = (w[:, :c1].contiguous(memory_format=torch.channels_last),
                               w[:, c1:].contiguous(memory_format=torch.channels_last))
        return self._sc_w[key]

    def _attention(self, att, x):
        """diffusers Attention block on a channels-last tensor: K5 without SiLU lands directly in the
        (B, HW, C) layout the projections want; q/k/v are one GEMM; to_out's bias rides with the residual."""
        B, C, H, W = x.shape
        y = groupnorm_silu_nhwc(x, att.group_norm, silu=False)
        h = y.permute(0, 2, 3, 1).reshape(B, H * W, C)
        wqkv, bqkv = self._qkv[id(att)]
        qkv = linear_tc(h, wqkv, bqkv) if _use_linear_tc(h, wqkv) else F.linear(h, wqkv, bqkv)      # (B, HW, 3C)
        if C // att.heads == 8 and H * W <= 64:
            o = attention_small(qkv, C)                                     # K7
        else:
            q, k, v = qkv.split(C, dim=-1)

            def split(t):
                return t.reshape(B, H * W, att.heads, C // att.heads).transpose(1, 2)
            o = F.scaled_dot_product_attention(split(q), split(k), split(v)).transpose(1, 2).reshape(B, H * W, C)
        wo = att.to_out[0].weight
        o = linear_tc(o, wo) if _use_linear_tc(o, wo) else F.linear(o, wo, None)
        o = o.reshape(B, H, W, C).permute(0, 3, 1, 2)                      # channels-last view
        return add_bias_residual_nhwc(x, o, att.to_out[0].bias)

    def _down(self, block, h, temb_act, h_bias=None):
        """Returns (h, skips, bias still owed to h): the downsampler's convolution runs without its bias; the next block's
        first resnet and the skip's consumer absorb it (like conv_in's)."""
        skips = []
        for i, resnet in enumerate(block.resnets):
            h = self._resnet(resnet, h, temb_act, x_bias=h_bias if i == 0 else None)
            if block.attentions is not None:
                h = self._attention(block.attentions[i], h)
            skips.append((h, None))
        if block.downsamplers is not None:
            conv = block.downsamplers[0].conv
            h = F.conv2d(h, conv.weight, None, stride=2, padding=1)
            skips.append((h, conv.bias))
            return h, skips, conv.bias
        return h, skips, None

    def _up(self, block, h, skips, temb_act, h_bias=None):
        """Returns (h, bias still owed to h): the upsampler's convolution runs without its bias, the next
        block's first resnet absorbs it (saves a full read+write of the upsampled activation)."""
        for i, resnet in enumerate(block.resnets):
            skip, skip_bias = skips.pop()
            h = self._resnet(resnet, h, temb_act, x2=skip, x_bias=h_bias if i == 0 else None, x2_bias=skip_bias)
            if block.attentions is not None:
                h = self._attention(block.attentions[i], h)
        if block.upsamplers is not None:
            conv = block.upsamplers[0].conv
            h = F.conv2d(upsample2x_nhwc(h), conv.weight, None, padding=1)
            return h, conv.bias
        return h, None

    kernels_per_forward = None      # K5 + K6 launches of one forward (set by the first call)
    supports_uniform_timestep = True

    @torch.no_grad()
    def forward(self, sample, timestep, return_dict=True, uniform_timestep=False):
        """``uniform_timestep``: the caller guarantees that every sample of the batch carries the SAME timestep (the
        samplers do: iadb_bn.py:306 builds ``tt`` from one t, ddim_diffusers.py:679 passes a scalar).  The whole
        time-embedding path -- sinusoid, two linears, SiLU and all 30 ``time_emb_proj`` projections -- then runs for
        ONE row (five matrix-vector products instead of fp32 SIMT GEMMs over the batch) and K5 broadcasts the row."""
        m = self.m
        launches_before = LAUNCHES
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.float32 if isinstance(t, float) else torch.int64, device=sample.device)
            uniform_timestep = True
        elif t.dim() == 0:
            t = t[None].to(sample.device)
            uniform_timestep = True
        if uniform_timestep:
            t = t[:1]
        else:
            t = t * torch.ones(sample.shape[0], dtype=t.dtype, device=t.device)
        emb = timestep_embedding(t, m.time_proj_dim)
        temb_act = F.silu(m.time_embedding(emb))
        temb_act = torch.addmm(self.temb_b, temb_act, self.temb_w.t())      # (B or 1, sum of Cout): all projections

        # conv_in's bias is owed to h: the first resnet and the last skip consumer absorb it
        h = conv_in3x3_nhwc(sample.float(), self._conv_in_w) if sample.is_cuda else None       # K11: NCHW state -> NHWC activation
        if h is None:
            h = F.conv2d(sample.float().contiguous(memory_format=torch.channels_last), m.conv_in.weight, None, padding=1)
        skips = [(h, m.conv_in.bias)]
        owed = m.conv_in.bias
        for block in m.down_blocks:
            h, s, owed = self._down(block, h, temb_act, h_bias=owed)
            skips.extend(s)
        h = self._resnet(m.mid_block.resnets[0], h, temb_act)
        h = self._attention(m.mid_block.attentions[0], h)
        h = self._resnet(m.mid_block.resnets[1], h, temb_act)
        owed = None
        for block in m.up_blocks:
            h, owed = self._up(block, h, skips, temb_act, h_bias=owed)
        h = m.conv_out(groupnorm_silu_nhwc(h, m.conv_norm_out, add_bc=owed))
        # left in channels-last memory: K2 (bndm_iadb_step_sched_dnhwc_f32) consumes it in place; any
        # other consumer sees an ordinary (B, C, H, W) tensor
        if self.kernels_per_forward is None:
            self.kernels_per_forward = LAUNCHES - launches_before
        if not return_dict:
            return (h,)
        return UNet2DOutput(sample=h)


def fuse_unet(model):
    """Returns the channels-last / fused-normalisation evaluator of ``model`` (a copy of its weights)."""
    return FusedUNet2D(model)
