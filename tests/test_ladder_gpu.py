"""Parity ladder step 4 and the fused-UNet isolation tests.

* GPU product vs the CPU REFERENCE end to end: tests/golden/ladder_e2e_tiny_unet.npz was written by executing the
  unmodified reference (get_noise_v2 -> utils.sample_iadb, fp32, CPU) around a small real UNet
  (oracle/make_golden.py::ladder_e2e); the GPU side replays it through K1, the stock / fused UNet and K2 with TF32
  off.  Tolerance: the north-star rtol with a 10x wider atol -- convolutions round differently on the two devices.
* fused vs stock UNet with TF32 off: the fusion itself (K5..K8, folded biases, split shortcuts) adds fp32 round-off
  only; with TF32 on the difference is cuDNN's choice of TF32 algorithm per layout and is bounded separately.
"""
import numpy as np
import pytest
import torch

import bndm_b200 as bb
from conftest import load_golden
from oracle.make_golden import state_sha, tiny_unet

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
E2E_RTOL, E2E_ATOL = 1e-4, 1e-4


@pytest.fixture
def no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.fixture(scope="module")
def L_dev(L_np):
    return torch.from_numpy(L_np).to(DEV)


@pytest.mark.parametrize("unet", ["stock", "fused"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_gpu_product_vs_cpu_reference_end_to_end(unet, use_graph, no_tf32, L_dev, L_sha):
    g = load_golden("ladder_e2e_tiny_unet")
    assert str(g["L_sha"]) == L_sha
    model = tiny_unet()
    if state_sha(model) != str(g["model_sha"]):
        pytest.skip("torch's CPU initialisation differs from the build that wrote the fixture")
    model = model.to(DEV)
    if unet == "fused":
        from bndm_b200.fused_unet import fuse_unet
        model = fuse_unet(model)
    white = torch.from_numpy(g["white"]).to(DEV)
    gamma = torch.from_numpy(g["gamma"]).to(DEV)
    x0 = bb.get_noise_v2(DEV, white, L_dev, gamma, None, "gaussianBN", "test", True)[0]
    np.testing.assert_allclose(x0.cpu().numpy(), g["x0"], rtol=1e-4, atol=1e-5)                # K1 alone: north-star tolerance
    with torch.no_grad():
        d = model(x0, torch.full((2,), 1.0, device=DEV), return_dict=False)[0]
    np.testing.assert_allclose(d.contiguous().cpu().numpy(), g["d_first"], rtol=E2E_RTOL, atol=E2E_ATOL)    # one UNet forward
    T = int(g["nb_step"])
    x, x_all, _ = bb.sample_iadb(model, x0, T, "sigmoid", tuple(g["scheduler_params"]), 6, "gaussianBN", "test", use_graph=use_graph)
    np.testing.assert_allclose(x.cpu().numpy(), g["x"], rtol=E2E_RTOL, atol=E2E_ATOL)
    assert len(x_all) == g["snaps"].shape[0]
    np.testing.assert_allclose(torch.stack(x_all).cpu().numpy(), g["snaps"], rtol=E2E_RTOL, atol=E2E_ATOL)


def _pair(which):
    from bndm_b200.fused_unet import fuse_unet
    from bndm_b200.unet import get_latent_model, get_model
    torch.manual_seed(0)
    if which == "cat_res64":
        model, x = get_model(3, 6, 64).to(DEV).eval(), torch.randn(4, 3, 64, 64, device=DEV)
    elif which == "cat_res64_b16":         # 16 x 16 tokens = 256 rows: the attention linears take K9 (3xTF32 tcgen05)
        model, x = get_model(3, 6, 64).to(DEV).eval(), torch.randn(16, 3, 64, 64, device=DEV)
    elif which == "cat_res128":
        model, x = get_model(3, 6, 128).to(DEV).eval(), torch.randn(2, 3, 128, 128, device=DEV)
    else:
        model, x = get_latent_model(512, 8).to(DEV).eval(), torch.randn(3, 4, 64, 64, device=DEV)
    t = torch.linspace(0.95, 0.004, x.shape[0], device=DEV)
    return model, fuse_unet(model), x, t


@pytest.mark.parametrize("which", ["cat_res64", "cat_res128", "latent512"])
def test_fused_unet_equals_stock_without_tf32(which, no_tf32):
    """TF32 off on both sides: what is left is the fusion's own fp32 round-off (other summation order inside the
    normalisation, biases added at another place) -- bounded at 1e-5 of the output scale."""
    model, fused, x, t = _pair(which)
    with torch.no_grad():
        want = model(x, t, return_dict=False)[0]
        got = fused(x, t, return_dict=False)[0]
    scale = max(1.0, want.abs().max().item())
    err = (got - want).abs().max().item()
    assert err <= 1e-5 * scale, (err, scale)


@pytest.mark.parametrize("which", ["cat_res64", "cat_res64_b16", "latent512"])
def test_fused_unet_vs_stock_with_tf32(which):
    """torch's default (TF32 convolutions allowed, as the reference runs): the two evaluations use different cuDNN TF32
    kernels (NHWC vs NCHW), each within TF32's 2^-11 operand rounding of the fp32 result."""
    model, fused, x, t = _pair(which)
    with torch.no_grad():
        want = model(x, t, return_dict=False)[0]
        got = fused(x, t, return_dict=False)[0]
    scale = max(1.0, want.abs().max().item())
    assert (got - want).abs().max().item() <= 5e-3 * scale


@pytest.mark.parametrize("which", ["cat_res64", "latent512"])
def test_uniform_timestep_path_equals_the_per_sample_path(which, no_tf32):
    """One timestep for the whole batch (what the reference's loops pass): the time-embedding path evaluated for ONE row and
    broadcast by K5 gives the per-sample evaluation's result to fp32 round-off of five small matrix products."""
    model, fused, x, _ = _pair(which)
    t = torch.full((x.shape[0],), 0.372, device=DEV)
    with torch.no_grad():
        a = fused(x, t, return_dict=False)[0]
        b = fused(x, t, return_dict=False, uniform_timestep=True)[0]
        c = fused(x, 0.372).sample                                    # python scalar: uniform by construction
        want = model(x, t, return_dict=False)[0]
    scale = max(1.0, want.abs().max().item())
    assert (a - b).abs().max().item() <= 1e-5 * scale and torch.equal(b, c)
    assert (b - want).abs().max().item() <= 1e-5 * scale
