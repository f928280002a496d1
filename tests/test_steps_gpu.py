"""GPU parity: the step kernels (K2 IADB, K3 DDIM, K4 uint8) and the samplers built on them,
through the C ABI, against the oracle / golden vectors.  Element-wise fp32 work with the
reference's association => bit-exact."""
import numpy as np
import pytest
import torch

import bndm_b200 as bb
from bndm_b200 import sampler as bs
from bndm_b200.ddim import sample_ddim
from conftest import ATOL, RTOL, SAMPLER_GOLDENS, assert_golden, load_golden
from oracle import sampler as osam
from oracle.toy import ToyCond, ToyEps

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("B,C,H,two", [(64, 3, 64, True), (5, 3, 16, True), (3, 4, 64, True), (7, 3, 64, False),
                                       (2, 3, 5, True), (1, 1, 1, False), (16, 4, 64, False), (32, 3, 128, True)])
def test_iadb_step_bit_exact(B, C, H, two):
    x = torch.randn(B, C, H, H, device=DEV)
    d = torch.randn(B, 2 * C if two else C, H, H, device=DEV)
    da, dg = torch.rand(B, device=DEV) * 0.01, torch.rand(B, device=DEV) * 0.01
    want = osam.iadb_update(x, d, da, dg, "gaussianBN", d.shape[1])
    got = bb.iadb_step(x, d, da, dg if two else None)
    assert torch.equal(got, want)
    # in place
    x2 = x.clone()
    bb.iadb_step(x2, d, da, dg if two else None, out=x2)
    assert torch.equal(x2, want)
    # CPU oracle gives the same bits (IEEE mul/add)
    want_cpu = osam.iadb_update(x.cpu(), d.cpu(), da.cpu(), dg.cpu(), "gaussianBN", d.shape[1])
    assert torch.equal(got.cpu(), want_cpu)


def test_iadb_step_rejects_bad_channel_count():
    x = torch.randn(2, 3, 8, 8, device=DEV)
    with pytest.raises(NotImplementedError):
        bb.iadb_step(x, torch.randn(2, 5, 8, 8, device=DEV), torch.ones(2, device=DEV), torch.ones(2, device=DEV))


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("name", SAMPLER_GOLDENS)
def test_sample_iadb_matches_reference_golden(name, use_graph):
    g = load_golden(name)
    oc, nt, T = int(g["out_channel"]), str(g["noise_type"]), int(g["nb_step"])
    x0 = torch.from_numpy(g["x0"].copy()).to(DEV)
    x, x_all, secs = bb.sample_iadb(ToyEps(oc), x0, T, "sigmoid", tuple(g["scheduler_params"]), oc, nt, "test",
                                    use_graph=use_graph)
    assert torch.equal(x0.cpu(), torch.from_numpy(g["x0"])), "x0 must not be modified"
    assert len(x_all) == int(g["n_snaps"]) and np.isfinite(secs)
    # bit-exact against the oracle evaluated on this host (same torch CPU schedule arithmetic) ...
    want, wall, _ = osam.sample_iadb_utils(ToyEps(oc), torch.from_numpy(g["x0"].copy()), T, "sigmoid",
                                           tuple(g["scheduler_params"]), oc, nt, "test")
    assert torch.equal(x.cpu(), want)
    for a, b in zip(x_all, wall):
        assert torch.equal(a.cpu(), b)
    # ... and against the vectors the real reference produced on the authoring host
    assert_golden(x.cpu().numpy(), g["x"], name)
    for i, idx in enumerate(g["snap_idx"]):
        assert_golden(x_all[int(idx)].cpu().numpy(), g["snaps"][i], name)


def test_sample_iadb_opt_signature_and_train_mode():
    x0 = torch.randn(3, 3, 16, 16, device=DEV)
    bs.opt.noise_type, bs.opt.out_channel, bs.opt.train_or_test = "gaussianBN", 6, "test"
    bs.opt.scheduler_alpha, bs.opt.scheduler_gamma = "linear", "sigmoid"
    x, x_all, _ = bb.sample_iadb(ToyEps(6), x0, 250, (1000.0, 0.0, 3.0))
    want, wall, _ = osam.sample_iadb_opt(ToyEps(6), x0.cpu(), 250, (1000.0, 0.0, 3.0), osam.make_opt())
    assert torch.equal(x.cpu(), want) and len(x_all) == len(wall) == 11
    for a, b in zip(x_all, wall):
        assert torch.equal(a.cpu(), b)
    bs.opt.train_or_test = "train"
    y = bb.sample_iadb(ToyEps(6), x0, 50, (0.2, 0.0, 3.0))
    assert torch.is_tensor(y)
    assert torch.equal(y.cpu(), osam.sample_iadb_opt(ToyEps(6), x0.cpu(), 50, (0.2, 0.0, 3.0),
                                                     osam.make_opt(train_or_test="train")))
    bs.opt.train_or_test = "test"
    bs.opt.noise_type = "bogus"
    with pytest.raises(NotImplementedError):
        bb.sample_iadb(ToyEps(6), x0, 5, (0.2, 0.0, 3.0))
    bs.opt.noise_type = "gaussianBN"


def test_sample_iadb_conditional():
    x0 = torch.randn(2, 3, 16, 16, device=DEV)
    xc = torch.randn(2, 3, 16, 16, device=DEV)
    bs.opt.noise_type, bs.opt.out_channel, bs.opt.train_or_test = "gaussianBN", 6, "test"
    x, x_all = bb.sample_iadb_conditional(ToyCond(6), x0, xc, 100, (0.2, 0.0, 3.0))
    want, wall = osam.sample_iadb_conditional(ToyCond(6), x0.cpu(), xc.cpu(), 100, (0.2, 0.0, 3.0), osam.make_opt())
    assert torch.equal(x.cpu(), want) and len(x_all) == len(wall)


@pytest.mark.parametrize("oc,nt", [(8, "gaussianBN"), (4, "gaussianBN"), (4, "gaussian")])
def test_latent_scheduler_and_loop(oc, nt):
    x = torch.randn(3, 4, 16, 16, device=DEV)
    d = torch.randn(3, oc, 16, 16, device=DEV)
    sch = bb.IADBScheduler(noise_type=nt, out_channels=oc)
    with pytest.raises(ValueError):
        sch.step(d, 3, x)
    sch.set_timesteps(250)
    got = sch.step(d, 17, x)
    want = osam.iadb_scheduler_step(d.cpu(), 17, x.cpu(), 250, nt, oc)
    assert torch.equal(got.cpu(), want)
    for g_ in (False, True):
        y = bb.sample_latent_iadb(ToyEps(oc), x, 50, nt, oc, use_graph=g_)
        assert torch.equal(y.cpu(), osam.latent_loop(ToyEps(oc), x.cpu(), 50, nt, oc))


@pytest.mark.parametrize("eta", [0.0, 1.0, 0.3])
@pytest.mark.parametrize("n", [100, 37])
def test_ddim_step_and_loop_vs_oracle(eta, n):
    tables = osam.DDIMTables()
    tables.set_timesteps(n)
    sch = bb.DDIMScheduler()
    sch.set_timesteps(n)
    x = torch.randn(4, 3, 16, 16, device=DEV) * 2
    eps = torch.randn(4, 3, 16, 16, device=DEV)
    vn = torch.randn(4, 3, 16, 16, device=DEV)
    for t in (int(sch.timesteps[0]), int(sch.timesteps[n // 2]), 0):
        got = sch.step(eps, t, x, eta=eta, variance_noise=vn).prev_sample
        want = osam.ddim_step(tables, eps.cpu(), t, x.cpu(), eta, vn.cpu())
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=RTOL, atol=ATOL)
        assert (got.cpu() - want).abs().max().item() <= 2e-6          # in practice <= 1-2 ulp
    noise_bank = torch.randn(n, 4, 3, 16, 16)
    fn_gpu = lambda i, t, xx: noise_bank[i].to(DEV)
    fn_cpu = lambda i, t, xx: noise_bank[i]
    got = sample_ddim(ToyEps(3), x, n, eta=eta, noise_fn=fn_gpu)
    want = osam.ddim_loop(ToyEps(3), x.cpu(), n, eta=eta, noise_fn=fn_cpu)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=RTOL, atol=ATOL)
    if eta == 0:
        gg = sample_ddim(ToyEps(3), x, n, use_graph=True)
        assert torch.equal(gg, got)


def test_ddim_with_time_varying_blue_noise(L_np):
    """BASELINE config 3: eta > 0, variance noise = get_noise_v2('gaussianBN', gamma(t))."""
    from oracle import noise as on
    L = torch.from_numpy(L_np).to(DEV)
    n, B = 10, 2
    x = torch.randn(B, 3, 64, 64, device=DEV)
    bank = torch.randn(n, B, 3, 64, 64)
    gam = torch.linspace(1, 0, n)

    def fn_gpu(i, t, xx):
        g = torch.full((B,), float(gam[i]), device=DEV)
        return bb.get_noise_v2(DEV, bank[i].to(DEV), L, g, t, "gaussianBN", "test", True)[0]

    def fn_cpu(i, t, xx):
        g = np.full((B,), float(gam[i]), np.float32)
        return torch.from_numpy(on.get_noise_np(bank[i].numpy(), L_np, g, "gaussianBN", "test", True)[0])
    got = sample_ddim(ToyEps(3), x, n, eta=1.0, noise_fn=fn_gpu)
    want = osam.ddim_loop(ToyEps(3), x.cpu(), n, eta=1.0, noise_fn=fn_cpu)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=RTOL, atol=ATOL)


def test_to_uint8_nhwc():
    from bndm_b200.io import to_uint8_nhwc
    x = torch.randn(3, 3, 16, 16, device=DEV) * 1.5
    got = to_uint8_nhwc(x)
    want = ((x / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1) * 255).round().to(torch.uint8)   # ddim_diffusers.py:687-688
    assert got.dtype == torch.uint8 and torch.equal(got, want)


def test_full_size_unet_sampling_graph_equals_eager_short():
    """Config-2 shaped run (64^2, out_channel 6, real UNet) for a few steps: the CUDA-graph path
    must reproduce the eager path exactly (same kernels, same order)."""
    from bndm_b200.unet import get_model
    torch.manual_seed(0)
    model = get_model(3, 6, 64).to(DEV).eval()
    x0 = torch.randn(2, 3, 64, 64, device=DEV)
    a = bb.sample_iadb(model, x0, 4, "sigmoid", (1000.0, 0.0, 3.0), 6, "gaussianBN", "train")
    b = bb.sample_iadb(model, x0, 4, "sigmoid", (1000.0, 0.0, 3.0), 6, "gaussianBN", "train", use_graph=True)
    assert torch.isfinite(a).all()
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-4, atol=1e-5)
    # and the oracle loop with the SAME module on the same device (parity ladder step 3)
    c = osam.sample_iadb_utils(model, x0, 4, "sigmoid", (1000.0, 0.0, 3.0), 6, "gaussianBN", "train")
    np.testing.assert_allclose(a.cpu().numpy(), c.cpu().numpy(), rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ K5: fused GroupNorm (+adds) + SiLU, NHWC
@pytest.mark.parametrize("B,C,H", [(3, 128, 64), (2, 256, 16), (2, 384, 8), (1, 512, 4), (2, 1024, 2), (2, 768, 8), (5, 128, 32),
                                   (2, 384, 64), (2, 256, 64), (1, 128, 128), (2, 256, 128), (1, 128, 256)])
@pytest.mark.parametrize("mode", ["plain", "add_bc", "res_sum", "nosilu"])
def test_groupnorm_silu_nhwc_matches_torch(B, C, H, mode):
    from bndm_b200.fused_unet import groupnorm_silu_nhwc
    torch.manual_seed(B * 1000 + C + H)
    x = (torch.randn(B, C, H, H, device=DEV) * 1.7 + 0.6).contiguous(memory_format=torch.channels_last)
    norm = torch.nn.GroupNorm(32, C, eps=1e-5).to(DEV)
    with torch.no_grad():
        norm.weight.copy_(torch.randn(C)); norm.bias.copy_(torch.randn(C))
    wide = torch.randn(B, 2 * C + 64, device=DEV)
    add_bc = wide[:, 32:32 + C] if mode == "add_bc" else None          # a column slice of a wider matrix (row stride != C)
    res = torch.randn_like(x) if mode == "res_sum" else None
    out = groupnorm_silu_nhwc(x, norm, add_bc=add_bc, res=res, want_sum=(mode == "res_sum"), silu=(mode != "nosilu"))
    s = x if res is None else x + res
    if add_bc is not None:
        s = s + add_bc[:, :, None, None]
    with torch.no_grad():
        want = norm(s)
        want64 = torch.nn.functional.group_norm(s.double(), 32, norm.weight.double(), norm.bias.double(), 1e-5)
        if mode != "nosilu":
            want, want64 = torch.nn.functional.silu(want), torch.nn.functional.silu(want64)
    if mode == "res_sum":
        y, ssum = out
        assert torch.equal(ssum, s)
    else:
        y = out
    assert y.is_contiguous(memory_format=torch.channels_last)
    err_ours = (y.double() - want64).abs().max().item()
    err_torch = (want.double() - want64).abs().max().item()
    assert err_ours <= max(2.0 * err_torch, 5e-6), (err_ours, err_torch)
    y2 = groupnorm_silu_nhwc(x, norm, add_bc=add_bc, res=res, silu=(mode != "nosilu"))
    assert torch.equal(y2, y), "fixed-order statistics must be bit-reproducible"


@pytest.mark.parametrize("B,C1,C2,H", [(2, 128, 128, 64), (2, 256, 128, 32), (3, 512, 512, 4), (2, 512, 256, 8), (2, 128, 128, 16),
                                       (2, 128, 128, 128), (1, 256, 128, 128)])
def test_groupnorm_two_sources_equals_cat(B, C1, C2, H):
    from bndm_b200.fused_unet import groupnorm_silu_nhwc
    a = torch.randn(B, C1, H, H, device=DEV).contiguous(memory_format=torch.channels_last)
    b = (torch.randn(B, C2, H, H, device=DEV) * 2 - 0.3).contiguous(memory_format=torch.channels_last)
    norm = torch.nn.GroupNorm(32, C1 + C2).to(DEV)
    with torch.no_grad():
        norm.weight.copy_(torch.randn(C1 + C2)); norm.bias.copy_(torch.randn(C1 + C2))
    tb = torch.randn(B, C1 + C2, device=DEV)
    got = groupnorm_silu_nhwc(a, norm, add_bc=tb, x2=b)
    want = groupnorm_silu_nhwc(torch.cat([a, b], 1), norm, add_bc=tb)
    assert torch.equal(got, want)


# ---- K9: the attention blocks' fp32 linears as a 3xTF32 tcgen05 GEMM (csrc/linear_tc.cu) ----
@pytest.mark.parametrize("M,N,K", [(1024, 1536, 512), (1024, 512, 512), (256, 1536, 512), (300, 516, 96), (128, 128, 32), (4096, 768, 256)])
@pytest.mark.parametrize("with_bias", [True, False])
def test_linear_tc_is_fp32_grade(M, N, K, with_bias):
    """out = a w^T (+ bias): the error against an fp64 product stays at the level of torch's fp32 GEMM (what the reference
    runs) and orders of magnitude under a single TF32 pass."""
    from bndm_b200.fused_unet import linear_tc
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV) * 1.3 + 0.2
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    b = torch.randn(N, device=DEV) if with_bias else None
    got = linear_tc(a, w, b)
    want64 = torch.nn.functional.linear(a.double(), w.double(), None if b is None else b.double())
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        ref32 = torch.nn.functional.linear(a, w, b)
        torch.backends.cuda.matmul.allow_tf32 = True
        tf32 = torch.nn.functional.linear(a, w, b)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    scale = want64.abs().max().item()
    err = (got.double() - want64).abs().max().item()
    err32 = (ref32.double() - want64).abs().max().item()
    err_tf32 = (tf32.double() - want64).abs().max().item()
    assert err <= max(4.0 * err32, 2e-6 * scale), (err, err32, err_tf32, scale)
    assert err * 20 < err_tf32 or err_tf32 < 1e-5 * scale, (err, err_tf32)
    assert torch.equal(linear_tc(a, w, b), got)
    # 3-D input (batch, tokens, features), as the attention block passes it
    if M % 16 == 0:
        assert torch.equal(linear_tc(a.view(M // 16, 16, K), w, b).reshape(M, N), got)


# ---- K10: 1x1 shortcut convolution of [x | x2] + residual + biases in one tcgen05 kernel (csrc/shortcut_tc.cu) ----
@pytest.mark.parametrize("B,C1,C2,N,H", [(2, 128, 128, 128, 64), (2, 256, 128, 128, 32), (3, 512, 512, 512, 4), (2, 512, 256, 256, 8),
                                         (2, 128, 0, 256, 16), (1, 128, 128, 128, 5), (5, 256, 256, 256, 16), (1, 96, 32, 132, 7)])
def test_shortcut_residual_matches_the_tf32_convolution(B, C1, C2, N, H):
    from bndm_b200.fused_unet import shortcut_residual_nhwc
    torch.manual_seed(B + C1 + C2 + N + H)
    cl = torch.channels_last
    x = torch.randn(B, C1, H, H, device=DEV).contiguous(memory_format=cl)
    x2 = torch.randn(B, C2, H, H, device=DEV).contiguous(memory_format=cl) if C2 else None
    w = (torch.randn(N, C1 + C2, 1, 1, device=DEV) / (C1 + C2) ** 0.5).contiguous(memory_format=cl)
    h2 = torch.randn(B, N, H, H, device=DEV).contiguous(memory_format=cl)
    bias = torch.randn(N, device=DEV)
    got = shortcut_residual_nhwc(x, x2, w, h2, bias)
    assert got.shape == h2.shape and got.is_contiguous(memory_format=cl)
    xc = x if x2 is None else torch.cat([x, x2], 1)
    want64 = torch.nn.functional.conv2d(xc.double(), w.double()) + h2.double() + bias.double()[None, :, None, None]
    old = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = True
        lib = torch.nn.functional.conv2d(xc.contiguous(memory_format=cl), w) + h2 + bias[None, :, None, None]
    finally:
        torch.backends.cudnn.allow_tf32 = old
    err = (got.double() - want64).abs().max().item()
    err_lib = (lib.double() - want64).abs().max().item()
    # TF32 inputs, fp32 accumulation: the same class of error as the library's TF32 convolution (which may have picked an
    # fp32 kernel for a small shape, hence the absolute floor at TF32's 2^-11 input rounding)
    assert err <= max(3.0 * err_lib, 4e-3), (err, err_lib)
    assert torch.equal(shortcut_residual_nhwc(x, x2, w, h2, bias), got)
    # in place on the residual
    h2c = h2.clone(memory_format=torch.preserve_format)
    from bndm_b200 import _lib
    M = B * H * H
    rc = _lib.load().bndm_shortcut_residual_tf32(_lib.ptr(x), _lib.ptr(x2), C1, C2, _lib.ptr(w.reshape(N, C1 + C2).contiguous()), _lib.ptr(h2c),
                                                 _lib.ptr(bias), _lib.ptr(h2c), M, N, _lib.current_stream(x.device))
    _lib.check(rc, "bndm_shortcut_residual_tf32")
    assert torch.equal(h2c, got)


# ---- K11: conv_in (3x3, <= 4 input channels) from the NCHW state to the channels-last activation (csrc/conv_in.cu) ----
@pytest.mark.parametrize("B,Cin,H,W,Cout", [(2, 3, 64, 64, 128), (1, 4, 32, 32, 128), (3, 3, 5, 7, 128), (2, 3, 16, 50, 64), (1, 3, 128, 128, 128),
                                            (2, 1, 9, 33, 32)])
def test_conv_in3x3_matches_conv2d(B, Cin, H, W, Cout):
    from bndm_b200.fused_unet import conv_in3x3_nhwc
    torch.manual_seed(B + Cin + H + W + Cout)
    x = torch.randn(B, Cin, H, W, device=DEV)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV) / (9 * Cin) ** 0.5
    got = conv_in3x3_nhwc(x, w)
    assert got is not None and got.shape == (B, Cout, H, W) and got.is_contiguous(memory_format=torch.channels_last)
    want = torch.nn.functional.conv2d(x.double(), w.double(), padding=1)
    assert (got.double() - want).abs().max().item() <= 1e-5
    assert torch.equal(conv_in3x3_nhwc(x, w), got)
    assert conv_in3x3_nhwc(torch.randn(1, 8, 4, 4, device=DEV), torch.randn(128, 8, 3, 3, device=DEV)) is None      # not a conv_in shape


def test_add_bias_residual_nhwc_is_bit_exact():
    from bndm_b200.fused_unet import add_bias_residual_nhwc
    a = torch.randn(3, 128, 16, 16, device=DEV).contiguous(memory_format=torch.channels_last)
    b = torch.randn(3, 128, 16, 16, device=DEV).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(128, device=DEV)
    got = add_bias_residual_nhwc(a, b, bias)
    assert got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, a + (b + bias[None, :, None, None]))
    bias_a = torch.randn(128, device=DEV)
    got = add_bias_residual_nhwc(a, b, bias, bias_a=bias_a)
    assert torch.equal(got, (a + bias_a[None, :, None, None]) + (b + bias[None, :, None, None]))
    a2 = torch.randn_like(a)
    got = add_bias_residual_nhwc(a, b, bias, bias_a=bias_a, a2=a2)
    assert torch.equal(got, ((a + a2) + bias_a[None, :, None, None]) + (b + bias[None, :, None, None]))


@pytest.mark.parametrize("B,T,C", [(3, 16, 512), (2, 4, 512), (5, 64, 64), (1, 1, 8)])
def test_attention_small_matches_sdpa(B, T, C):
    from bndm_b200.fused_unet import attention_small
    qkv = torch.randn(B, T, 3 * C, device=DEV)
    got = attention_small(qkv, C)
    q, k, v = (t.reshape(B, T, C // 8, 8).transpose(1, 2) for t in qkv.split(C, dim=-1))
    want = torch.nn.functional.scaled_dot_product_attention(q.double(), k.double(), v.double()).transpose(1, 2).reshape(B, T, C)
    np.testing.assert_allclose(got.cpu().numpy(), want.float().cpu().numpy(), rtol=1e-5, atol=2e-6)


def test_fused_unet_matches_plain_unet():
    from bndm_b200.fused_unet import fuse_unet
    from bndm_b200.unet import get_model
    torch.manual_seed(0)
    model = get_model(3, 6, 64).to(DEV).eval()
    fused = fuse_unet(model)
    x = torch.randn(4, 3, 64, 64, device=DEV)
    t = torch.tensor([0.9, 0.5, 0.25, 0.004], device=DEV)
    with torch.no_grad():
        want = model(x, t, return_dict=False)[0]
        got = fused(x, t, return_dict=False)[0]
        again = fused(x, t).sample
    assert got.shape == want.shape
    scale = want.abs().max().item()
    # same cuDNN TF32 convolutions (possibly other algorithms in NHWC) + fp32 round-off of the norms
    assert (got - want).abs().max().item() < 5e-3 * scale, ((got - want).abs().max().item(), scale)
    assert torch.equal(got, again)


def test_sample_iadb_with_fused_unet_in_a_cuda_graph():
    """K5/K6/K7 are stream-ordered and allocation-free: the [fused UNet -> K2] step captures into one
    CUDA graph and replays bit-identically to the eager loop with the same module."""
    from bndm_b200.fused_unet import fuse_unet
    from bndm_b200.unet import get_latent_model
    torch.manual_seed(0)
    model = fuse_unet(get_latent_model(256, 8).to(DEV).eval())
    z = torch.randn(2, 4, 32, 32, device=DEV)
    eager = bb.sample_latent_iadb(model, z, 5, "gaussianBN", 8, use_graph=False)
    graphed = bb.sample_latent_iadb(model, z, 5, "gaussianBN", 8, use_graph=True)
    assert torch.equal(eager, graphed)
    ref = osam.latent_loop(model, z, 5, "gaussianBN", 8)          # the oracle's loop around the same module
    np.testing.assert_allclose(graphed.cpu().numpy(), ref.cpu().numpy(), rtol=RTOL, atol=ATOL)


def test_fused_unet_res128_matches_plain():
    """cfg 4's UNet (7 blocks, 128x128): exercises the cluster-of-16 GroupNorm path."""
    from bndm_b200.fused_unet import fuse_unet
    from bndm_b200.unet import get_model
    torch.manual_seed(1)
    model = get_model(3, 6, 128).to(DEV).eval()
    fused = fuse_unet(model)
    x = torch.randn(2, 3, 128, 128, device=DEV)
    t = torch.tensor([0.7, 0.1], device=DEV)
    with torch.no_grad():
        want = model(x, t, return_dict=False)[0]
        got = fused(x, t, return_dict=False)[0]
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() < 5e-3 * scale


@pytest.mark.parametrize("B,C,Cd,H", [(5, 3, 6, 64), (3, 4, 8, 32), (2, 3, 3, 16)])
def test_sched_step_consumes_channels_last_output_in_place(B, C, Cd, H):
    """K2 reading the UNet output in NHWC memory == K2 reading the NCHW copy, bit for bit."""
    from bndm_b200.schedules import iadb_table
    table, first_t = iadb_table(6, batch=B)
    x = torch.randn(B, C, H, H, device=DEV)
    d = torch.randn(B, Cd, H, H, device=DEV)
    d_cl = d.contiguous(memory_format=torch.channels_last)
    assert not d_cl.is_contiguous()
    outs = []
    for dd in (d, d_cl):
        st = bs.IadbStepper(table, first_t, B, DEV)
        xx = x.clone()
        for _ in range(3):
            st.step_(xx, dd)
        outs.append((xx, st.t_vec.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("B,C,H,W", [(3, 128, 32, 32), (2, 512, 2, 2), (1, 256, 5, 7)])
def test_upsample2x_nhwc_is_exact(B, C, H, W):
    from bndm_b200.fused_unet import upsample2x_nhwc
    x = torch.randn(B, C, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
    got = upsample2x_nhwc(x)
    assert got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, torch.nn.functional.interpolate(x, scale_factor=2.0, mode="nearest"))


@pytest.mark.parametrize("use_graph", [False, True])
def test_sample_ddim_with_channels_last_model_output(use_graph):
    """A model that returns its prediction in channels-last memory (like FusedUNet2D) must give the
    same DDIM trajectory as the same model returning NCHW memory."""
    from types import SimpleNamespace

    class CL:
        def __init__(self, inner, channels_last):
            self.inner, self.cl = inner, channels_last

        def __call__(self, x, t, return_dict=True):
            y = self.inner(x, t).sample
            if self.cl:
                y = y.contiguous(memory_format=torch.channels_last)
                assert not y.is_contiguous()
            return SimpleNamespace(sample=y)
    x = torch.randn(3, 3, 16, 16, device=DEV)
    a = sample_ddim(CL(ToyEps(3), False), x, 10, use_graph=use_graph)
    b = sample_ddim(CL(ToyEps(3), True), x, 10, use_graph=use_graph)
    assert torch.equal(a, b)


def test_graph_samplers_are_cached_and_stay_correct_across_calls():
    """sample_ddim / sample_latent_iadb keep their captured step per (model, shape, schedule): the second call
    replays the cached graph on new inputs and must equal the eager loop for those inputs."""
    from bndm_b200 import ddim as bd
    from bndm_b200.unet import get_latent_model
    eps_model = ToyEps(3)
    xs = [torch.randn(3, 3, 16, 16, device=DEV) for _ in range(3)]
    graphs = set()
    for x in xs:
        got = sample_ddim(eps_model, x, 10, use_graph=True)
        want = sample_ddim(eps_model, x, 10, use_graph=False)
        assert torch.equal(got, want)
        graphs.add(id([e for k, e in bd._graph_cache.items() if k[0] == id(eps_model)][0]["graph"]))
    assert len(graphs) == 1 and len(bd._graph_cache) <= 4, "one captured graph serves all three calls"
    torch.manual_seed(0)
    model = get_latent_model(256, 8).to(DEV).eval()
    for seed in range(2):
        z = torch.randn(2, 4, 32, 32, device=DEV, generator=torch.Generator(device=DEV).manual_seed(seed))
        got = bb.sample_latent_iadb(model, z, 3, "gaussianBN", 8, use_graph=True)
        want = bb.sample_latent_iadb(model, z, 3, "gaussianBN", 8, use_graph=False)
        assert torch.equal(got, want)
    assert len(bs._sampler_cache) <= 4
