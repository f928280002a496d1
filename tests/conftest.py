import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def L_np():
    from bndm_b200.synth import hashed_tril
    return hashed_tril(seed=0)


@pytest.fixture(scope="session")
def L_sha(L_np):
    return hashlib.sha256(L_np.tobytes()).hexdigest()[:16]


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


NOISE_GOLDENS = ["noise64_bn_b4c3_cfg1", "noise64_gbn_b2c4", "noise64_rn_b2c3_draw", "noise32_bn_b2c4",
                 "noise32_bn_b1c4_draw", "noise128_bn_b2c3", "noise128_bn_b1c3_draw", "noise128_gauss_test_b2c3"]
SAMPLER_GOLDENS = ["sampler_bn_oc6", "sampler_bn_oc6_tau02", "sampler_gauss_oc3", "sampler_gbn_oc3_T1000"]

RTOL, ATOL = 1e-4, 1e-5          # BASELINE.json north_star: fp32 tolerance of the floating-point path


def assert_golden(got, want, what=""):
    """Golden vectors were produced by the reference on the authoring container's CPU.  Data
    movement and IEEE add/mul chains reproduce bit for bit anywhere; schedules go through
    torch's CPU sigmoid/cos (vectorised libm, last-ulp differences between AVX2 / AVX-512
    hosts), so on another host the comparison falls back to a tolerance 10x tighter than the
    north-star one.  Bit-exactness on the SAME host is asserted separately against the oracle."""
    got, want = np.asarray(got), np.asarray(want)
    if np.array_equal(got, want):
        return
    np.testing.assert_allclose(got, want, rtol=RTOL / 10, atol=ATOL / 10, err_msg=what)
