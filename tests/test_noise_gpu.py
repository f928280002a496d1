"""GPU parity: get_noise_v2 (libbndm_b200.so through the C ABI) against the CPU oracle, the
reference-generated golden vectors and size-independent properties."""
import numpy as np
import pytest
import torch

import bndm_b200 as bb
from bndm_b200 import _lib
from conftest import ATOL, NOISE_GOLDENS, RTOL, load_golden
from oracle import noise as on

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GEMMS = ["tc", "simt", "gemv"]
FLAG = {"auto": _lib.GEMM_AUTO, "tc": _lib.GEMM_TC, "simt": _lib.GEMM_SIMT, "gemv": _lib.GEMM_GEMV}


def _skip_wide_gemv(gemm, n_cols):
    """K1g covers the GEMV regime only (<= 16 GEMM columns); wider calls are K1b's."""
    if gemm == "gemv" and n_cols > 16:
        pytest.skip("K1g handles <= 16 columns")


def _cols(res, bs, C):
    return bs * C * (4 if res == 128 else 1)


@pytest.fixture(scope="module")
def L_dev(L_np):
    return torch.from_numpy(L_np).to(DEV)


def _np(t):
    return t.detach().cpu().numpy()


def test_native_library_is_loaded():
    lib = _lib.load()
    assert lib.bndm_version() == 2
    with torch.cuda.device(DEV):
        assert lib.bndm_device_is_sm100() == 1, "tests expect a B200 (sm_100)"
    with open("/proc/self/maps") as f:
        assert "libbndm_b200.so" in f.read()


@pytest.mark.parametrize("gemm", GEMMS)
@pytest.mark.parametrize("name", NOISE_GOLDENS)
def test_golden(name, gemm, L_dev):
    g = load_golden(name)
    nt, inplace, tt = str(g["noise_type"]), bool(g["inplace"]), str(g["train_or_test"])
    if nt != "gaussian":
        _skip_wide_gemv(gemm, _cols(g["x"].shape[-1], g["x"].shape[0], g["x"].shape[1]))
    x = torch.from_numpy(g["x"].copy()).to(DEV)
    gamma = torch.from_numpy(g["gamma"]).to(DEV)
    if inplace or nt == "gaussian":
        out, bn, wn = bb.get_noise_v2(DEV, x, L_dev, gamma, None, nt, tt, inplace, gemm=gemm)
    else:
        # replay the reference's draw: the C-ABI call takes the white field explicitly
        out, bn, wn = _call_with_draw(L_dev, torch.from_numpy(g["draw"]).to(DEV), gamma, x.shape, nt, gemm)
    np.testing.assert_allclose(_np(out), g["out"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(_np(bn), g["bn"], rtol=RTOL, atol=ATOL)
    assert np.array_equal(_np(wn), g["wn"])


def _call_with_draw(L_dev, draw, gamma, shape, noise_type, gemm):
    bs, C, res, _ = shape
    h = bb.prepare_L(L_dev)
    out = torch.empty(shape, device=DEV)
    bn = torch.empty(shape, device=DEV)
    wn = torch.empty(shape, device=DEV)
    g = None if noise_type == "GBN" else gamma
    rc = _lib.load().bndm_get_noise_f32(h._h, _lib.ptr(draw), _lib.ptr(g), _lib.ptr(out), _lib.ptr(bn), _lib.ptr(wn),
                                        bs, C, res, _lib.SRC_DRAW | FLAG[gemm],
                                        _lib.current_stream(DEV))
    _lib.check(rc, "bndm_get_noise_f32")
    return out, bn, wn


# BASELINE shard shapes: cfg 1 (64, 4, 3); cfg 2 (64, 64, 3) = 2 x 96 columns; cfg 4 per GPU (128, 32, 3) = 384 columns
# = 3 x 128 through the 128^2 output map; cfg 5 per GPU (64, 16, 4) = 64 columns.  (32, 64, 4) and (32, 50, 3) /
# (128, 16, 3) and (128, 9, 4) are the multi-column-block cases of the 32^2 / 128^2 branches.
CASES = [(64, 4, 3), (64, 1, 1), (64, 7, 4), (64, 64, 3), (32, 3, 4), (32, 1, 1), (128, 2, 3), (128, 1, 4),
         (128, 5, 3), (64, 100, 3), (128, 32, 3), (64, 16, 4), (32, 64, 4), (32, 50, 3), (128, 16, 3), (128, 9, 4),
         (64, 5, 3), (64, 4, 4), (128, 1, 3)]


@pytest.mark.parametrize("gemm", GEMMS)
@pytest.mark.parametrize("noise_type", ["gaussianBN", "GBN"])
@pytest.mark.parametrize("res,bs,C", CASES)
def test_vs_oracle_inplace(res, bs, C, noise_type, gemm, L_np, L_dev):
    _skip_wide_gemv(gemm, _cols(res, bs, C))
    if gemm == "simt" and _cols(res, bs, C) > 300:
        pytest.skip("the fp32 witness kernel is covered at smaller sizes")
    rng = np.random.default_rng(res + 7 * bs + C)
    x = rng.standard_normal((bs, C, res, res)).astype(np.float32)
    gamma = rng.random(bs).astype(np.float32)
    want = on.get_noise_np(x, L_np, gamma, noise_type, "train", True)
    xt = torch.from_numpy(x).to(DEV)
    got = bb.get_noise_v2(DEV, xt, L_dev, torch.from_numpy(gamma).to(DEV), None, noise_type, "train", True, gemm=gemm)
    assert torch.equal(xt.cpu(), torch.from_numpy(x)), "input must not be modified"
    for g_, w_, nm in zip(got, want, ("noise", "bn", "wn")):
        assert tuple(g_.shape) == w_.shape, nm
        if nm == "wn":
            assert np.array_equal(_np(g_), w_)
        else:
            np.testing.assert_allclose(_np(g_), w_, rtol=RTOL, atol=ATOL, err_msg=nm)
    if noise_type == "GBN":
        assert got[0] is got[1]                      # get_noise_recent.py:118: noise = noise_bn


@pytest.mark.parametrize("res,bs,C", [(64, 3, 3), (32, 2, 4), (128, 2, 3), (128, 32, 3), (32, 64, 4), (64, 16, 4), (128, 1, 3)])
def test_draw_path_uses_torch_rng_like_reference(res, bs, C, L_np, L_dev):
    x = torch.zeros(bs, C, res, res, device=DEV)
    gamma = torch.rand(bs, device=DEV)
    torch.manual_seed(11)
    got = bb.get_noise_v2(DEV, x, L_dev, gamma, None, "gaussianBN", "train", False)
    torch.manual_seed(11)                           # replay the draw exactly as the reference issues it
    if res == 64:
        draw = torch.randn_like(x)
    elif res == 32:
        draw = torch.randn_like(torch.zeros(bs, C, 64, 64, device=DEV))
    else:
        draw = torch.randn(bs * 4, C, 64, 64).float().to(DEV)
    want = on.get_noise_np(_np(x), L_np, _np(gamma), "gaussianBN", "train", False, _np(draw))
    np.testing.assert_allclose(_np(got[0]), want[0], rtol=RTOL, atol=ATOL)
    assert np.array_equal(_np(got[2]), want[2])


def test_gaussian_passthrough(L_dev):
    x64 = torch.randn(2, 3, 64, 64, device=DEV)
    a, b, c = bb.get_noise_v2(DEV, x64, L_dev, None, None, "gaussian", "train", True)
    assert a is x64 and b is x64 and c is x64
    x = torch.randn(3, 3, 128, 128, device=DEV)
    got = bb.get_noise_v2(DEV, x, L_dev, None, None, "gaussian", "test", True)[0]
    want = on.get_noise_np(_np(x), None, None, "gaussian", "test", True)[0]
    assert np.array_equal(_np(got), want)
    torch.manual_seed(3)
    tr = bb.get_noise_v2(DEV, x, L_dev, None, None, "gaussian", "train", False)[0]
    torch.manual_seed(3)
    assert torch.equal(tr, torch.randn_like(x))


def test_errors_mirror_reference(L_dev):
    with pytest.raises(NotImplementedError):
        bb.get_noise_v2(DEV, torch.randn(1, 3, 16, 16, device=DEV), L_dev, torch.ones(1, device=DEV), None, "gaussianBN")
    with pytest.raises(NotImplementedError):
        bb.get_noise_v2(DEV, torch.randn(1, 3, 256, 256, device=DEV), L_dev, None, None, "gaussian")
    with pytest.raises(ValueError):
        bb.get_noise_v2(DEV, torch.randn(2, 3, 64, 64, device=DEV), L_dev, torch.ones(3, device=DEV), None, "gaussianBN",
                        "train", True)
    with pytest.raises(NotImplementedError):       # bndm_prepare_L: L must be (4096, 4096)
        bb.prepare_L(torch.eye(1024, device=DEV))


# ------------------------------------------------------------------ properties (any size)
@pytest.mark.parametrize("gemm", GEMMS)
def test_gamma_one_is_white_and_gamma_zero_is_blue(gemm, L_dev):
    x = torch.randn(5, 3, 64, 64, device=DEV)
    one = torch.ones(5, device=DEV)
    out, bn, wn = bb.get_noise_v2(DEV, x, L_dev, one, None, "gaussianBN", "train", True, gemm=gemm)
    assert torch.equal(wn, x)
    assert torch.equal(out, bn * 0.0 + x)           # bn*(1-1) + wn*1, exactly
    out0, bn0, _ = bb.get_noise_v2(DEV, x, L_dev, one * 0, None, "gaussianBN", "train", True, gemm=gemm)
    assert torch.equal(out0, bn0 + x * 0.0)


@pytest.mark.parametrize("gemm", GEMMS)
def test_identity_L_returns_white(gemm):
    L = torch.eye(4096, device=DEV)
    x = torch.randn(5, 3, 64, 64, device=DEV)
    bn = bb.get_noise_v2(DEV, x, L, None, None, "GBN", "train", True, gemm=gemm)[1]
    # 3xTF32 reconstructs hi+lo exactly up to the dropped 2^-22 term
    np.testing.assert_allclose(_np(bn), _np(x), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("gemm", GEMMS)
def test_linearity_and_determinism(gemm, L_dev):
    a = torch.randn(4, 3, 64, 64, device=DEV)
    b = torch.randn(4, 3, 64, 64, device=DEV)
    f = lambda t: bb.get_noise_v2(DEV, t, L_dev, None, None, "GBN", "train", True, gemm=gemm)[1]
    np.testing.assert_allclose(_np(f(a + b)), _np(f(a) + f(b)), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(_np(f(a * 2.0)), _np(f(a) * 2.0), rtol=1e-6, atol=1e-6)
    assert torch.equal(f(a), f(a)), "split-K combine must be run-to-run reproducible"


@pytest.mark.parametrize("gemm", GEMMS)
def test_triangular_skip_equals_dense(gemm, L_np, L_dev):
    x = torch.randn(3, 3, 64, 64, device=DEV)
    h = bb.prepare_L(L_dev)
    assert h.lower_triangular
    outs = []
    for extra in (0, _lib.FORCE_DENSE):
        o = torch.empty_like(x)
        rc = _lib.load().bndm_get_noise_f32(h._h, _lib.ptr(x), None, _lib.ptr(o), None, None, 3, 3, 64,
                                            _lib.SRC_IMAGE | FLAG[gemm] | extra,
                                            _lib.current_stream(DEV))
        _lib.check(rc, "get_noise")
        outs.append(o)
    np.testing.assert_allclose(_np(outs[0]), _np(outs[1]), rtol=RTOL, atol=ATOL)
    # a genuinely dense L is detected and handled
    rng = np.random.default_rng(3)
    Ld = (rng.standard_normal((4096, 4096)) / 64).astype(np.float32)
    hd = bb.prepare_L(torch.from_numpy(Ld).to(DEV))
    assert not hd.lower_triangular
    got = bb.get_noise_v2(DEV, x, hd, None, None, "GBN", "train", True, gemm=gemm)[1]
    want = on.get_noise_np(_np(x), Ld, None, "GBN", "train", True)[1]
    np.testing.assert_allclose(_np(got), want, rtol=RTOL, atol=ATOL)


def test_gemv_matches_tc_and_simt_closely(L_dev):
    x = torch.randn(4, 3, 64, 64, device=DEV)
    g = torch.rand(4, device=DEV)
    r = {k: bb.get_noise_v2(DEV, x, L_dev, g, None, "gaussianBN", "train", True, gemm=k)[0] for k in ("gemv", "tc", "simt", "auto")}
    assert torch.equal(r["auto"], r["gemv"]), "12 columns: the default rule picks K1g (<= 16 columns)"
    assert (r["gemv"] - r["simt"]).abs().max().item() < 4e-6
    assert (r["gemv"] - r["tc"]).abs().max().item() < 4e-6


def test_forced_kernel_flags(L_dev):
    x = torch.randn(8, 3, 64, 64, device=DEV)                      # 24 columns: beyond K1g
    with pytest.raises(NotImplementedError):
        bb.get_noise_v2(DEV, x, L_dev, None, None, "GBN", "train", True, gemm="gemv")
    a = bb.get_noise_v2(DEV, x, L_dev, None, None, "GBN", "train", True, gemm="auto")[0]
    b = bb.get_noise_v2(DEV, x, L_dev, None, None, "GBN", "train", True, gemm="tc")[0]
    assert torch.equal(a, b), "24 columns: the default rule picks K1b"


def test_tc_matches_simt_closely(L_dev):
    x = torch.randn(64, 3, 64, 64, device=DEV)
    g = torch.rand(64, device=DEV)
    a = bb.get_noise_v2(DEV, x, L_dev, g, None, "gaussianBN", "train", True, gemm="tc")[0]
    b = bb.get_noise_v2(DEV, x, L_dev, g, None, "gaussianBN", "train", True, gemm="simt")[0]
    err = (a - b).abs().max().item()
    assert err < 4e-6, err


def test_blue_spectrum_is_high_pass():
    """The figure script's check (scripts/fig_main_3_4_inset_10_supp_1_2.py:31-36,121-122): the
    power spectrum of L.z for a blue-noise L has (much) less energy at low frequencies."""
    from bndm_b200.synth import blue_noise_L
    L = torch.from_numpy(blue_noise_L()).to(DEV)
    x = torch.randn(32, 3, 64, 64, device=DEV)
    bn = bb.get_noise_v2(DEV, x, L, None, None, "GBN", "train", True)[1]
    spec = torch.fft.fft2(bn).abs().pow(2).mean(dim=(0, 1))
    fy = torch.fft.fftfreq(64, device=DEV)
    fr = (fy[:, None] ** 2 + fy[None, :] ** 2).sqrt()
    low, high = spec[(fr > 0) & (fr < 0.1)].mean(), spec[fr > 0.35].mean()
    assert low < 0.2 * high, (low.item(), high.item())
    np.testing.assert_allclose(bn.var().item(), 1.0, rtol=0.05)         # unit-diagonal covariance


def test_cuda_graph_capture_of_get_noise(L_dev):
    x = torch.randn(4, 3, 64, 64, device=DEV)
    g = torch.rand(4, device=DEV)
    h = bb.prepare_L(L_dev)
    h.reserve(12)
    eager = bb.get_noise_v2(DEV, x, h, g, None, "gaussianBN", "train", True)[0].clone()
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(graph, stream=s):
        out = bb.get_noise_v2(DEV, x, h, g, None, "gaussianBN", "train", True)[0]
    x.copy_(torch.randn_like(x))
    graph.replay()
    torch.cuda.synchronize()
    again = bb.get_noise_v2(DEV, x, h, g, None, "gaussianBN", "train", True)[0]
    assert torch.equal(out, again) and not torch.equal(out, eager)


# ------------------------------------------------------------------ contraction variants
@pytest.fixture
def policy():
    """Forces the tcgen05 kernel's variants through bndm_debug_set_policy; restores the default rule."""
    lib = _lib.load()
    yield lambda fused, raw: lib.bndm_debug_set_policy(fused, raw)
    lib.bndm_debug_set_policy(-1, -1)


@pytest.mark.parametrize("res,bs,C,noise_type,inplace", [(64, 4, 3, "gaussianBN", True), (64, 4, 4, "gaussianBN", True),
                                                         (64, 1, 3, "GBN", True), (32, 3, 4, "gaussianBN", True),
                                                         (32, 4, 4, "gaussianBN", False), (128, 1, 3, "gaussianBN", True),
                                                         (128, 1, 4, "gaussianBN", False), (64, 2, 4, "gaussianRN", False)])
def test_gemv_instances_match_oracle(res, bs, C, noise_type, inplace, L_np, L_dev):
    """K1g at every column-count instantiation (4, 8, 12, 16), every branch, through the C ABI with its own handle;
    run-to-run bit-identical."""
    h = bb.CovMatL(L_dev)
    rng = np.random.default_rng(77 + res + bs + C)
    x = rng.standard_normal((bs, C, res, res)).astype(np.float32)
    gamma = rng.random(bs).astype(np.float32)
    draw = None if inplace else rng.standard_normal((bs * (4 if res == 128 else 1), C, 64, 64)).astype(np.float32)
    want = on.get_noise_np(x, L_np, gamma, noise_type, "train", inplace, draw)
    shape = (bs, C, res, res)
    outs = []
    for _ in range(2):
        o = [torch.empty(shape, device=DEV) for _ in range(3)]
        src = torch.from_numpy(x if inplace else draw).to(DEV)
        g = None if noise_type == "GBN" else torch.from_numpy(gamma).to(DEV)
        rc = _lib.load().bndm_get_noise_f32(h._h, _lib.ptr(src), _lib.ptr(g), _lib.ptr(o[0]), _lib.ptr(o[1]), _lib.ptr(o[2]),
                                            bs, C, res, (_lib.SRC_IMAGE if inplace else _lib.SRC_DRAW) | _lib.GEMM_GEMV,
                                            _lib.current_stream(DEV))
        _lib.check(rc, "bndm_get_noise_f32")
        outs.append(o)
    for a_, b_, w_, nm in zip(outs[0], outs[1], want, ("noise", "bn", "wn")):
        assert torch.equal(a_, b_), nm
        if nm == "wn":
            assert np.array_equal(_np(a_), w_)
        else:
            np.testing.assert_allclose(_np(a_), w_, rtol=RTOL, atol=ATOL, err_msg=nm)
    h.close()


@pytest.mark.parametrize("res,bs,C,inplace", [(128, 8, 3, True), (128, 6, 3, False), (64, 8, 3, True), (32, 8, 4, True),
                                              (32, 6, 4, False), (128, 3, 4, True)])
def test_sharded_call_equals_rows_of_the_unsharded_call(res, bs, C, inplace, L_dev):
    """SURVEY 8e seed-parity rule incl. the 128^2 inplace branch, whose (b', k') = divmod(k*bs + b, 4) mixing uses the
    GLOBAL batch (get_noise_recent.py:131-146): every shard of the global field reproduces the same rows of the full
    call -- the white field bit for bit; the blue field to a few ulp (the contraction's instance and split-K cut
    depend on the column count, so different shard sizes add the same products in another order), and bit for bit
    whenever the shard runs the same kernel instance (same shard size, or the batch-invariant K1g)."""
    x = torch.randn(bs, C, res, res, device=DEV)
    gamma = torch.rand(bs, device=DEV)
    torch.manual_seed(5)
    full = bb.get_noise_v2(DEV, x, L_dev, gamma, None, "gaussianBN", "test", inplace)
    parts = {}
    for lo, hi in ((0, bs // 2), (bs // 2, bs), (1, 2), (0, bs)):
        torch.manual_seed(5)
        part = parts[(lo, hi)] = bb.get_noise_v2(DEV, x, L_dev, gamma, None, "gaussianBN", "test", inplace, shard=(lo, hi))
        for f_, p_, nm in zip(full, part, ("noise", "bn", "wn")):
            assert p_.shape[0] == hi - lo
            if nm == "wn" or (lo, hi) == (0, bs):
                assert torch.equal(p_, f_[lo:hi]), (nm, lo, hi)
            else:
                np.testing.assert_allclose(_np(p_), _np(f_[lo:hi]), rtol=1e-5, atol=2e-6, err_msg=f"{nm} {lo}:{hi}")
    # the same global field cut at the same size on "another rank": bit-identical (this is what N-GPU == 1-GPU means)
    torch.manual_seed(5)
    again = bb.get_noise_v2(DEV, x.clone(), L_dev, gamma.clone(), None, "gaussianBN", "test", inplace, shard=(0, bs // 2))
    for a_, p_ in zip(again, parts[(0, bs // 2)]):
        assert torch.equal(a_, p_)
    if res == 128:          # the 'gaussian' test-mode pass-through has the same mixing (:50-56)
        fullg = bb.get_noise_v2(DEV, x, L_dev, None, None, "gaussian", "test", True)[0]
        partg = bb.get_noise_v2(DEV, x, L_dev, None, None, "gaussian", "test", True, shard=(1, bs - 1))[0]
        assert torch.equal(partg, fullg[1:bs - 1])


def test_gemv_is_batch_invariant(L_dev):
    """K1g adds a column's products in an order that does not depend on the other columns: a sample's result is the
    same bits whether it is computed alone or in a batch of 4."""
    x = torch.randn(4, 3, 64, 64, device=DEV)
    g = torch.rand(4, device=DEV)
    full = bb.get_noise_v2(DEV, x, L_dev, g, None, "gaussianBN", "train", True, gemm="gemv")
    for b in range(4):
        one = bb.get_noise_v2(DEV, x, L_dev, g, None, "gaussianBN", "train", True, gemm="gemv", shard=(b, b + 1))
        for f_, o_ in zip(full, one):
            assert torch.equal(o_, f_[b:b + 1])


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("raw", [0, 1])
@pytest.mark.parametrize("res,bs,C", [(64, 4, 3), (64, 20, 3), (64, 40, 3), (32, 5, 4), (128, 3, 3), (64, 50, 3)])
def test_contraction_variants_match_oracle(res, bs, C, fused, raw, policy, L_np, L_dev):
    """fused combine on/off x raw-L converter on/off (nb <= 64) give the oracle's result; the
    same variant run twice is bit-identical."""
    policy(fused, raw)
    rng = np.random.default_rng(1000 + res + bs)
    x = rng.standard_normal((bs, C, res, res)).astype(np.float32)
    gamma = rng.random(bs).astype(np.float32)
    want = on.get_noise_np(x, L_np, gamma, "gaussianBN", "train", True)
    xt, gt = torch.from_numpy(x).to(DEV), torch.from_numpy(gamma).to(DEV)
    got = bb.get_noise_v2(DEV, xt, L_dev, gt, None, "gaussianBN", "train", True, gemm="tc")
    again = bb.get_noise_v2(DEV, xt, L_dev, gt, None, "gaussianBN", "train", True, gemm="tc")
    for g_, a_, w_, nm in zip(got, again, want, ("noise", "bn", "wn")):
        assert torch.equal(g_, a_), nm
        if nm == "wn":
            assert np.array_equal(_np(g_), w_)
        else:
            np.testing.assert_allclose(_np(g_), w_, rtol=RTOL, atol=ATOL, err_msg=nm)


@pytest.mark.parametrize("bs", [4, 64])
def test_unit_variance_blue_L_accuracy(bs):
    """A real Cholesky factor (unit-diagonal covariance => unit-variance output): the 3xTF32
    contraction stays an order of magnitude inside atol against an fp64 product."""
    from bndm_b200.synth import blue_noise_L
    L = torch.from_numpy(blue_noise_L()).to(DEV)
    x = torch.randn(bs, 3, 64, 64, device=DEV)
    ref = (x.double().reshape(bs * 3, 4096) @ L.double().T).reshape(bs, 3, 64, 64)
    for gemm in GEMMS:
        if gemm == "gemv" and bs * 3 > 16:
            continue
        bn = bb.get_noise_v2(DEV, x, L, None, None, "GBN", "train", True, gemm=gemm)[1]
        err = (bn.double() - ref).abs().max().item()
        assert err < (4e-6 if gemm in ("tc", "gemv") else ATOL), (gemm, err)


# ------------------------------------------------------------------ training-side fusion
@pytest.mark.parametrize("gemm", GEMMS)
@pytest.mark.parametrize("res,bs,C,noise_type", [(64, 6, 3, "gaussianBN"), (64, 64, 3, "gaussianBN"), (32, 4, 4, "gaussianBN"),
                                                 (128, 2, 3, "gaussianBN"), (64, 5, 3, "GBN"), (64, 4, 3, "gaussianBN"),
                                                 (128, 32, 3, "gaussianBN"), (32, 64, 4, "gaussianBN"), (64, 16, 4, "gaussianBN"),
                                                 (128, 1, 3, "gaussianBN")])
def test_get_noise_train_equals_the_torch_sequence(res, bs, C, noise_type, gemm, L_dev):
    _skip_wide_gemv(gemm, _cols(res, bs, C))
    if gemm == "simt" and _cols(res, bs, C) > 300:
        pytest.skip("the fp32 witness kernel is covered at smaller sizes")
    """x_alpha / tar1 / tar2 from the fused epilogue == the reference's expressions (iadb_bn.py:915,949-950)
    applied to this library's own (x0, bn, wn) for the same draw, bit for bit."""
    g = torch.Generator(device="cpu").manual_seed(res + bs)
    x1 = (torch.rand(bs, C, res, res, generator=g) * 2 - 1).to(DEV)
    draw = torch.randn(bs * (4 if res == 128 else 1), C, 64, 64, generator=g).to(DEV)
    gamma = torch.rand(bs, generator=g).to(DEV)
    alpha = torch.rand(bs, generator=g).to(DEV)
    alpha_prev = torch.rand(bs, generator=g).to(DEV)
    x0, bn, wn = _call_with_draw(L_dev, draw, gamma, x1.shape, noise_type, gemm)
    xa, t1, t2, x0f = bb.get_noise_train(DEV, x1, L_dev, gamma, alpha, alpha_prev, noise_type, draw=draw, want_x0=True, gemm=gemm)
    a4 = alpha.view(-1, 1, 1, 1)
    assert torch.equal(x0f, x0)
    assert torch.equal(xa, a4 * x0 + (1 - a4) * x1)
    assert torch.equal(t1, x1 - x0)
    if noise_type == "GBN":
        assert t2 is None
    else:
        assert torch.equal(t2, alpha_prev.view(-1, 1, 1, 1) * (bn - wn))
