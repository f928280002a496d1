"""GPU tests of the pieces around the hot path: output post-processing, L construction on the device (N4),
and the robustness contracts of the stepper / handle / conditional sampler."""
import numpy as np
import pytest
import torch

import bndm_b200 as bb
import bndm_b200.sampler as bs
from bndm_b200 import _lib
from conftest import ATOL, RTOL
from oracle import sampler as osam
from oracle.toy import ToyEps

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def L_dev(L_np):
    return torch.from_numpy(L_np).to(DEV)


# ------------------------------------------------------------------ N3: the IADB driver's PNG conversion on the device
@pytest.mark.parametrize("C,H", [(3, 64), (3, 128), (4, 32), (1, 17)])
def test_snapshots_uint8_equal_the_reference_conversion(C, H):
    from bndm_b200.io import iadb_snapshot_uint8, iadb_snapshots_uint8
    torch.manual_seed(C * 100 + H)
    x = torch.randn(6, C, H, H, device=DEV) * 1.3 + 0.2
    final = [False, False, True, False, True, False]
    got = iadb_snapshots_uint8(x, final)
    assert got.dtype == torch.uint8 and got.shape == (6, H, H, C)
    for n in range(6):
        want = iadb_snapshot_uint8(x[n], final[n])              # the reference's torch / numpy expressions (iadb_bn.py:796-802)
        assert np.array_equal(got[n].cpu().numpy(), want), n
    assert torch.equal(iadb_snapshots_uint8(x, True)[1], torch.from_numpy(iadb_snapshot_uint8(x[1], True)).to(DEV))
    assert torch.equal(iadb_snapshots_uint8(x, False)[2], torch.from_numpy(iadb_snapshot_uint8(x[2], False)).to(DEV))


# ------------------------------------------------------------------ N4: covariance -> Cholesky factor on the device
def test_device_cholesky_matches_numpy_fp64_at_full_size():
    """n = 4096 (the size cov_mat_L must have): float64 potrf on the GPU vs numpy's float64 Cholesky of the same
    covariance, and the factor drives get_noise_v2 (unit-variance blue field)."""
    from bndm_b200.synth import blue_noise_sigma, cholesky_L, red_noise_sigma
    for sigma in (blue_noise_sigma(), red_noise_sigma()):
        want = np.linalg.cholesky(sigma).astype(np.float32)
        got = cholesky_L(torch.from_numpy(sigma).to(DEV))
        assert got.dtype == torch.float32 and got.is_cuda and torch.equal(got, torch.tril(got))
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=2e-6)
    x = torch.randn(16, 3, 64, 64, device=DEV)
    bn = bb.get_noise_v2(DEV, x, got, None, None, "GBN", "train", True)[1]
    np.testing.assert_allclose(bn.var().item(), 1.0, rtol=0.05)


def test_empirical_covariance_on_device():
    from bndm_b200.synth import cholesky_L, empirical_covariance
    g = torch.Generator(device="cpu").manual_seed(0)
    fields = torch.randn(3000, 16, 16, generator=g).to(DEV)
    cov = empirical_covariance(fields)
    assert cov.is_cuda and cov.dtype == torch.float64 and cov.shape == (256, 256)
    want = np.cov(fields.reshape(3000, -1).double().cpu().numpy(), rowvar=False)
    np.testing.assert_allclose(cov.cpu().numpy(), want, rtol=1e-9, atol=1e-12)
    L = cholesky_L(cov, jitter=1e-6)
    np.testing.assert_allclose((L.double() @ L.double().T).cpu().numpy(), want + 1e-6 * np.eye(256), rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ stepper contracts
def test_stepper_refuses_to_run_past_its_schedule_and_never_reads_past_the_table():
    from bndm_b200.schedules import iadb_table
    B, T = 3, 4
    table, first_t = iadb_table(T, batch=B)
    st = bs.IadbStepper(table, first_t, B, DEV, expect_channels=6)
    x = torch.randn(B, 3, 16, 16, device=DEV)
    d = torch.randn(B, 6, 16, 16, device=DEV)
    for _ in range(T):
        st.step_(x, d)
    assert not st.overrun
    with pytest.raises(RuntimeError):
        st.step_(x, d)                                    # host-side guard
    # the device-side guard: drive the C ABI directly one launch too far -- the last row is re-used, the flag is set
    before = x.clone()
    rc = _lib.load().bndm_iadb_step_sched_f32(_lib.ptr(x), _lib.ptr(x), _lib.ptr(d), _lib.ptr(st.table), _lib.ptr(st.state),
                                              _lib.ptr(st.t_vec), B, 3, 256, 6, _lib.current_stream(DEV))
    _lib.check(rc, "step")
    assert st.overrun
    row = st.table[T - 1]
    want = (before + row[:, 0].view(-1, 1, 1, 1) * d[:, :3]) + row[:, 1].view(-1, 1, 1, 1) * d[:, 3:]
    assert torch.equal(x, want)
    st.reset()
    assert not st.overrun
    with pytest.raises(ValueError):                       # a model that emits the wrong channel count is not accepted
        st.step_(x, torch.randn(B, 3, 16, 16, device=DEV))
    st.reset()
    st.step_(x, d)
    with pytest.raises(RuntimeError):                     # the kernel variant (grid) is fixed within a run
        st.step_(x, d.contiguous(memory_format=torch.channels_last))


def test_first_timestep_is_per_sample_and_schedule_can_be_evaluated_on_the_gpu():
    from bndm_b200.schedules import iadb_table
    B, T, params = 5, 50, (1000.0, 0.0, 3.0)
    table, first = iadb_table(T, "sigmoid", "sigmoid", params, 2.0, batch=B)
    st = bs.IadbStepper(table, first, B, DEV)
    assert torch.equal(st.t_vec.cpu(), first)             # the (B,) vector alpha_start of the first step (iadb_bn.py:311)
    # evaluated on the sampling device, the table equals the oracle's loop evaluated there (a reference on a GPU box)
    tg, fg = iadb_table(T, "linear", "sigmoid", params, batch=B, device=DEV)
    for row, t in enumerate(reversed(range(T))):
        a_s, a_e, g_s, g_e = osam._coefficients(t, B, DEV, T, "linear", "sigmoid", params)
        assert torch.equal(tg[row, :, 0], (a_s - a_e).cpu()) and torch.equal(tg[row, :, 1], (g_s - g_e).cpu())
    x0 = torch.randn(B, 3, 16, 16, device=DEV)
    got = bb.sample_iadb(ToyEps(6), x0, T, "sigmoid", params, 6, "gaussianBN", "train", schedule_device=DEV)
    want = osam.sample_iadb_utils(ToyEps(6), x0, T, "sigmoid", params, 6, "gaussianBN", "train")
    assert torch.equal(got, want)


def test_conditional_sampler_reuses_one_graph_for_new_conditioning():
    """x_c lives in a static buffer of the sampler: a new batch's x_c (another tensor, or the same tensor updated in
    place) replays the SAME captured graph with the new values."""
    opt = bs.opt
    old = (opt.noise_type, opt.out_channel, opt.train_or_test, opt.nb_steps)
    opt.noise_type, opt.out_channel, opt.train_or_test, opt.nb_steps = "gaussianBN", 6, "train", 5
    try:
        from oracle.toy import ToyCond
        model = ToyCond(6)
        x0 = torch.randn(2, 3, 16, 16, device=DEV)
        keys_before = set(bs._sampler_cache)
        outs, wants = [], []
        xc = torch.randn(2, 3, 16, 16, device=DEV)
        for i in range(3):
            if i == 1:
                xc = torch.randn(2, 3, 16, 16, device=DEV)        # another tensor
            if i == 2:
                xc.mul_(-2.0)                                     # same tensor, new contents
            outs.append(bb.sample_iadb_conditional(model, x0, xc, 5, (1000.0, 0.0, 3.0), use_graph=True))
            wants.append(osam.sample_iadb_conditional(model, x0, xc, 5, (1000.0, 0.0, 3.0), osam.make_opt(
                noise_type="gaussianBN", out_channel=6, train_or_test="train", nb_steps=5)))
        assert len(set(bs._sampler_cache) - keys_before) == 1      # one capture served all three (the cache holds at most 4)
        for g_, w_ in zip(outs, wants):
            np.testing.assert_allclose(g_.cpu().numpy(), w_.cpu().numpy(), rtol=RTOL, atol=ATOL)
    finally:
        opt.noise_type, opt.out_channel, opt.train_or_test, opt.nb_steps = old


# ------------------------------------------------------------------ one handle, two streams
def test_one_handle_on_two_streams_is_serialised(L_dev, L_np):
    """A bndm_L handle owns one workspace: calls issued on different streams must not overlap on it.  Two streams hammer
    the same handle with different inputs (the pre-split tcgen05 path: pack -> zt -> partials); every result must be right."""
    from oracle import noise as on
    h = bb.prepare_L(L_dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    xs = [torch.randn(40, 3, 64, 64, device=DEV) for _ in range(2)]
    want = [on.get_noise_np(x.cpu().numpy(), L_np, None, "GBN", "train", True)[1] for x in xs]
    torch.cuda.synchronize()
    outs = [[], []]
    for it in range(6):
        for k, s in enumerate((s1, s2)):
            with torch.cuda.stream(s):
                outs[k].append(bb.get_noise_v2(DEV, xs[k], h, None, None, "GBN", "train", True, gemm="tc")[1])
    torch.cuda.synchronize()
    for k in range(2):
        for o in outs[k]:
            np.testing.assert_allclose(o.cpu().numpy(), want[k], rtol=RTOL, atol=ATOL)
    # a second stream while the first one is still being captured cannot be ordered: refused, not silently raced
    g = torch.cuda.CUDAGraph()
    s3 = torch.cuda.Stream()
    s3.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=s3):
        captured = bb.get_noise_v2(DEV, xs[0], h, None, None, "GBN", "train", True, gemm="tc")[1]
        with pytest.raises(bb.BndmError):
            with torch.cuda.stream(s1):
                bb.get_noise_v2(DEV, xs[1], h, None, None, "GBN", "train", True, gemm="tc")
    g.replay()
    torch.cuda.synchronize()
    np.testing.assert_allclose(captured.cpu().numpy(), want[0], rtol=RTOL, atol=ATOL)


# ------------------------------------------------------------------ N1: whole-loop graph, checkpoint files
def test_whole_loop_graph_equals_per_step_graph():
    """All T steps captured as ONE CUDA graph (K2 keeps the step index on the device) == T replays of the one-step graph
    == the eager loop, bit for bit; the second run replays the same graph with a new x0."""
    from bndm_b200.fused_unet import fuse_unet
    from bndm_b200.unet import get_latent_model
    torch.manual_seed(0)
    model = fuse_unet(get_latent_model(256, 8).to(DEV).eval())
    for seed in (1, 2):
        x0 = torch.randn(2, 4, 32, 32, generator=torch.Generator().manual_seed(seed)).to(DEV)
        args = (7, "sigmoid", (0.2, 0.0, 3.0), 8, "gaussianBN", "train")
        eager = bb.sample_iadb(model, x0, *args)
        step = bb.sample_iadb(model, x0, *args, use_graph=True)
        loop = bb.sample_iadb(model, x0, *args, use_graph="loop")
        assert torch.equal(eager, step) and torch.equal(step, loop)
    with pytest.raises(ValueError):
        bb.sample_iadb(model, x0, 7, "sigmoid", (0.2, 0.0, 3.0), 8, "gaussianBN", "test", use_graph="loop")


def test_checkpoint_files_with_diffusers_key_names_load(tmp_path):
    """iadb_bn.py:714 `model.load_state_dict(torch.load(model.ckpt))` and ddim_diffusers.py:642 `from_pretrained` (a
    safetensors file): a state dict under the diffusers parameter names round-trips through both file formats into a
    fresh model, and the fused evaluator built from the loaded model reproduces the original's output."""
    from bndm_b200.fused_unet import fuse_unet
    from bndm_b200.unet import get_latent_model
    torch.manual_seed(3)
    src = get_latent_model(256, 8).to(DEV).eval()
    sd = {k: v.detach().cpu() for k, v in src.state_dict().items()}
    for needle in ("conv_in.weight", "time_embedding.linear_1.weight", "down_blocks.0.resnets.0.norm1.weight",
                   "down_blocks.0.resnets.0.time_emb_proj.bias", "down_blocks.0.downsamplers.0.conv.weight",
                   "mid_block.attentions.0.to_q.weight", "mid_block.attentions.0.to_out.0.bias",
                   "up_blocks.0.resnets.0.conv_shortcut.weight", "up_blocks.0.upsamplers.0.conv.weight",
                   "conv_norm_out.weight", "conv_out.bias"):
        assert needle in sd, needle
    x = torch.randn(2, 4, 32, 32, device=DEV)
    t = torch.tensor([0.8, 0.3], device=DEV)
    with torch.no_grad():
        want = src(x, t, return_dict=False)[0]
    torch.save(sd, tmp_path / "model.ckpt")
    files = {"ckpt": lambda: torch.load(tmp_path / "model.ckpt", map_location="cpu")}
    try:
        from safetensors.torch import load_file, save_file
        save_file(sd, str(tmp_path / "diffusion_pytorch_model.safetensors"))
        files["safetensors"] = lambda: load_file(str(tmp_path / "diffusion_pytorch_model.safetensors"))
    except ImportError:
        pass
    for name, load in files.items():
        torch.manual_seed(99)                                   # a differently initialised model ...
        fresh = get_latent_model(256, 8).to(DEV).eval()
        missing, unexpected = fresh.load_state_dict(load(), strict=True)      # ... takes every key, none left over
        assert not missing and not unexpected, name
        with torch.no_grad():
            assert torch.equal(fresh(x, t, return_dict=False)[0], want), name
            got = fuse_unet(fresh)(x, t, return_dict=False)[0]
        assert (got - want).abs().max().item() <= 5e-3 * max(1.0, want.abs().max().item()), name


def test_prepare_L_cache_is_tied_to_live_memory(L_np):
    """The handle cache can never serve stale operand copies: an entry pins its tensor, in-place edits drop it, and a
    different matrix -- even one that would have been allocated at a recycled address -- gets its own handle."""
    from bndm_b200 import noise as bn
    from oracle import noise as on
    x = torch.randn(2, 3, 64, 64, device=DEV)
    La = torch.from_numpy(L_np).to(DEV)
    ha = bb.prepare_L(La)
    assert bb.prepare_L(La) is ha
    ptr = La.data_ptr()
    del La                                             # the caller drops its tensor: the cache entry keeps the memory pinned
    torch.cuda.empty_cache()
    Lb = torch.from_numpy(np.ascontiguousarray(np.tril(L_np.T * 0.5 + L_np))).to(DEV)       # another matrix
    assert Lb.data_ptr() != ptr
    hb = bb.prepare_L(Lb)
    assert hb is not ha
    want = on.get_noise_np(x.cpu().numpy(), Lb.cpu().numpy(), None, "GBN", "train", True)[1]
    np.testing.assert_allclose(bb.get_noise_v2(DEV, x, Lb, None, None, "GBN", "train", True)[1].cpu().numpy(), want, rtol=RTOL, atol=ATOL)
    Lb.mul_(2.0)                                       # in-place edit: the cached copies are stale, a new handle is built
    hc = bb.prepare_L(Lb)
    assert hc is not hb
    np.testing.assert_allclose(bb.get_noise_v2(DEV, x, Lb, None, None, "GBN", "train", True)[1].cpu().numpy(), 2 * want, rtol=RTOL, atol=ATOL)
    assert len(bn._handles) <= bn._MAX_HANDLES


@pytest.mark.parametrize("sub,bs", [("1", 4), ("2", 8)])
def test_env_selected_contraction_instances(sub, bs, tmp_path):
    """The k-stages-per-pipeline-stage knob BNDM_TC_SUB reaches two instances the default rule never picks
    (gemm_tc_kernel<16,true,1> and <32,true,2>); it is read once per process, so each runs in its own interpreter."""
    import os
    import subprocess
    import sys
    code = f"""
import numpy as np, torch, sys
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import bndm_b200 as bb
from bndm_b200.synth import hashed_tril
from oracle import noise as on
dev = torch.device('cuda:0')
L_np = hashed_tril(seed=0)
L = torch.from_numpy(L_np).to(dev)
rng = np.random.default_rng(5)
x = rng.standard_normal(({bs}, 3, 64, 64)).astype(np.float32)
g = rng.random({bs}).astype(np.float32)
want = on.get_noise_np(x, L_np, g, 'gaussianBN', 'train', True)
got = bb.get_noise_v2(dev, torch.from_numpy(x).to(dev), L, torch.from_numpy(g).to(dev), None, 'gaussianBN', 'train', True, gemm='tc')
again = bb.get_noise_v2(dev, torch.from_numpy(x).to(dev), L, torch.from_numpy(g).to(dev), None, 'gaussianBN', 'train', True, gemm='tc')
for a, b, w in zip(got, again, want):
    assert torch.equal(a, b)
    np.testing.assert_allclose(a.cpu().numpy(), w, rtol=1e-4, atol=1e-5)
print('ok')
"""
    env = dict(os.environ, BNDM_TC_SUB=sub, BNDM_TC_RAWL="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "ok" in out.stdout, (out.stdout + out.stderr)[-2000:]
