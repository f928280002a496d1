"""CPU: the oracle restatements against the golden vectors the REAL reference produced
(oracle/make_golden.py) -- this is what pins the oracle where /root/reference is absent."""
import numpy as np
import pytest
import torch

from conftest import ATOL, NOISE_GOLDENS, RTOL, SAMPLER_GOLDENS, assert_golden, load_golden
from oracle import noise as on
from oracle import sampler as osam
from oracle import schedules as osch
from oracle.toy import ToyEps


@pytest.mark.parametrize("name", NOISE_GOLDENS)
def test_noise_np_matches_golden(name, L_np, L_sha):
    g = load_golden(name)
    assert str(g["L_sha"]) == L_sha, "hashed_tril is not bit-reproducible on this machine"
    draw = g["draw"] if g["draw"].size else None
    out, bn, wn = on.get_noise_np(g["x"], L_np, g["gamma"], str(g["noise_type"]), str(g["train_or_test"]),
                                  bool(g["inplace"]), draw)
    np.testing.assert_allclose(out, g["out"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(bn, g["bn"], rtol=RTOL, atol=ATOL)
    assert np.array_equal(wn, g["wn"])            # white field: pure data movement -> bit exact


@pytest.mark.parametrize("name", NOISE_GOLDENS)
def test_noise_torch_matches_golden_bitwise(name, L_np):
    g = load_golden(name)
    torch.manual_seed(int(g["seed"]))
    out, bn, wn = on.get_noise_torch(torch.device("cpu"), torch.from_numpy(g["x"].copy()), torch.from_numpy(L_np),
                                     torch.from_numpy(g["gamma"]), None, str(g["noise_type"]),
                                     str(g["train_or_test"]), bool(g["inplace"]))
    # same ops as the reference; the GEMM may pick another blocking on another CPU -> tolerance
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(bn.numpy(), g["bn"], rtol=RTOL, atol=ATOL)
    assert np.array_equal(wn.numpy(), g["wn"])


def test_schedules_match_golden():
    g = load_golden("schedules")
    for key in g.files:
        parts = key.split("_")
        T = int(parts[-1][1:])
        x = torch.arange(0, T + 1).float()
        if parts[0] == "alpha":
            got = osch.alpha_schedule(x, "linear", T)
        else:
            kind = parts[1]
            p = (float(parts[2][3:]), float(parts[3][1:]), float(parts[4][1:]))
            got = osch.gamma_schedule(x, kind, p, T)
        assert_golden(got.numpy(), g[key], key)


@pytest.mark.parametrize("name", SAMPLER_GOLDENS)
def test_sampler_matches_golden_bitwise(name):
    g = load_golden(name)
    oc, nt, T = int(g["out_channel"]), str(g["noise_type"]), int(g["nb_step"])
    x, snaps, _ = osam.sample_iadb_utils(ToyEps(oc), torch.from_numpy(g["x0"].copy()), T, "sigmoid",
                                         tuple(g["scheduler_params"]), oc, nt, "test")
    assert len(snaps) == int(g["n_snaps"])
    assert_golden(x.numpy(), g["x"], name)
    for i, idx in enumerate(g["snap_idx"]):
        assert_golden(snaps[int(idx)].numpy(), g["snaps"][i], name)


def test_opt_variant_shares_arithmetic_with_utils():
    # iadb_bn.sample_iadb differs from utils.sample_iadb only in the opt global and the snapshot cadence
    x0 = torch.randn(2, 3, 8, 8)
    opt = osam.make_opt(noise_type="gaussianBN", out_channel=6)
    xa, sa, _ = osam.sample_iadb_opt(ToyEps(6), x0, 250, (1000.0, 0.0, 3.0), opt)
    xb, sb, _ = osam.sample_iadb_utils(ToyEps(6), x0, 250, "sigmoid", (1000.0, 0.0, 3.0), 6, "gaussianBN", "test")
    assert torch.equal(xa, xb)
    assert len(sa) == 11 and len(sb) == 250           # t = 249, 225, ..., 0  vs every step
    assert torch.equal(sa[-1], sb[-1])


def test_latent_step_is_python_float_update():
    x = torch.randn(2, 4, 8, 8)
    d = torch.randn(2, 8, 8, 8)
    y = osam.iadb_scheduler_step(d, 10, x, 250, "gaussianBN", 8)
    c = (11 / 250 - 10 / 250)
    assert torch.equal(y, x + c * d[:, :4] + c * d[:, 4:])
    with pytest.raises(ValueError):
        osam.iadb_scheduler_step(d, 10, x, None, "gaussianBN", 8)
    with pytest.raises(NotImplementedError):
        osam.iadb_scheduler_step(d, 10, x, 250, "GBN", 8)


def test_ddim_tables_and_step_properties():
    tb = osam.DDIMTables()
    tb.set_timesteps(100)
    assert list(tb.timesteps[:3]) == [990, 980, 970] and tb.timesteps[-1] == 0
    sa, sb, sap, sdir, sigma = tb.coefficients(0, 0.0)
    assert float(sap) == 1.0 and float(sdir) == 0.0 and float(sigma) == 0.0     # set_alpha_to_one
    x, eps = torch.randn(2, 3, 8, 8), torch.randn(2, 3, 8, 8)
    # last step with eta = 0 returns the clipped x0 prediction
    y = osam.ddim_step(tb, eps, 0, x, 0.0)
    assert torch.equal(y, ((x - sb * eps) / sa).clamp(-1, 1))
    # sqrt(abar)^2 + sqrt(1-abar)^2 = 1
    for t in (990, 500, 10):
        a, b, *_ = tb.coefficients(t)
        assert abs(float(a) ** 2 + float(b) ** 2 - 1) < 1e-6
