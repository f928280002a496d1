"""CPU, authoring container only: the oracle against the UNMODIFIED reference functions
imported from /root/reference (skipped where the reference tree is absent, e.g. the GPU box;
tests/test_oracle_golden.py covers the same ground there through committed fixtures)."""
import numpy as np
import pytest
import torch

from conftest import ATOL, RTOL
from oracle import noise as on
from oracle import sampler as osam
from oracle import schedules as osch
from oracle.ref_import import load_reference, reference_available
from oracle.toy import ToyEps

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return load_reference()


def _draw_shape(res, bs, C):
    return {32: (bs, C, 64, 64), 64: (bs, C, 64, 64), 128: (bs * 4, C, 64, 64)}[res]


@pytest.mark.parametrize("res,C,bs", [(64, 3, 4), (32, 4, 2), (128, 3, 2), (64, 4, 1)])
@pytest.mark.parametrize("noise_type", ["gaussianBN", "GBN", "gaussianRN"])
@pytest.mark.parametrize("inplace", [True, False])
def test_blue_branches(ref, L_np, res, C, bs, noise_type, inplace):
    get_noise_v2 = ref[0]
    rng = np.random.default_rng(res * 10 + C)
    x = rng.standard_normal((bs, C, res, res)).astype(np.float32)
    gamma = rng.random(bs).astype(np.float32)
    Lt = torch.from_numpy(L_np)
    torch.manual_seed(5)
    r = get_noise_v2(torch.device("cpu"), torch.from_numpy(x.copy()), Lt, torch.from_numpy(gamma), None, noise_type,
                     "train", inplace)
    torch.manual_seed(5)
    draw = None if inplace else torch.randn(*_draw_shape(res, bs, C)).numpy()
    o = on.get_noise_np(x, L_np, gamma, noise_type, "train", inplace, draw)
    np.testing.assert_allclose(o[0], r[0].numpy(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(o[1], r[1].numpy(), rtol=RTOL, atol=ATOL)
    assert np.array_equal(o[2], r[2].numpy())
    torch.manual_seed(5)
    t = on.get_noise_torch(torch.device("cpu"), torch.from_numpy(x.copy()), Lt, torch.from_numpy(gamma), None,
                           noise_type, "train", inplace)
    for a, b in zip(t, r):
        assert torch.equal(a, b)                      # same ops in the same order -> bit exact


@pytest.mark.parametrize("res", [64, 128])
@pytest.mark.parametrize("tt", ["train", "test"])
@pytest.mark.parametrize("inplace", [True, False])
def test_gaussian_passthrough(ref, res, tt, inplace):
    get_noise_v2 = ref[0]
    x = torch.randn(2, 3, res, res)
    torch.manual_seed(1)
    r = get_noise_v2(torch.device("cpu"), x.clone(), None, None, None, "gaussian", tt, inplace)
    torch.manual_seed(1)
    draw = torch.randn(2, 3, res, res).numpy()
    o = on.get_noise_np(x.numpy(), None, None, "gaussian", tt, inplace, draw)
    assert np.array_equal(o[0], r[0].numpy())


def test_unsupported_sizes_raise_like_reference(ref, L_np):
    get_noise_v2 = ref[0]
    x = torch.randn(1, 3, 16, 16)
    for nt in ("gaussian", "gaussianBN"):
        with pytest.raises(NotImplementedError):
            get_noise_v2(torch.device("cpu"), x, torch.from_numpy(L_np), torch.rand(1), None, nt, "train", True)
        with pytest.raises(NotImplementedError):
            on.get_noise_np(x.numpy(), L_np, np.ones(1, np.float32), nt, "train", True)


@pytest.mark.parametrize("T", [250, 1000, 100])
def test_schedules(ref, T):
    ru = ref[2]
    x = torch.arange(0, T + 1).float()
    assert torch.equal(osch.alpha_schedule(x, "linear", T), ru.get_scheduler(x, "linear", T))
    for kind, p in [("sigmoid", (1000, 0, 3)), ("sigmoid", (0.2, 0, 3)), ("cosine", (1, 0.2, 1)),
                    ("cosine", (2, 0.1, 0.9)), ("linear", (1, 0, 3))]:
        assert torch.equal(osch.gamma_schedule(x, kind, p, T), ru.get_scheduler_gamma(x, kind, p, T)), (kind, p)


@pytest.mark.parametrize("nt,oc", [("gaussianBN", 6), ("gaussianBN", 3), ("gaussian", 3), ("GBN", 3)])
def test_sampler(ref, nt, oc):
    ru = ref[2]
    x0 = torch.randn(3, 3, 8, 8)
    r = ru.sample_iadb(ToyEps(oc), x0, 50, "sigmoid", (1000, 0, 3), oc, nt, "test")
    o = osam.sample_iadb_utils(ToyEps(oc), x0, 50, "sigmoid", (1000, 0, 3), oc, nt, "test")
    assert torch.equal(o[0], r[0]) and len(o[1]) == len(r[1]) == 50
    for a, b in zip(o[1], r[1]):
        assert torch.equal(a, b)
    r2 = ru.sample_iadb(ToyEps(oc), x0, 20, "sigmoid", (0.2, 0, 3), oc, nt, "train")
    o2 = osam.sample_iadb_utils(ToyEps(oc), x0, 20, "sigmoid", (0.2, 0, 3), oc, nt, "train")
    assert torch.equal(o2, r2)
