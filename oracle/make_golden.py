"""Generate tests/golden/*.npz by EXECUTING THE REAL REFERENCE (authoring container only).

    python -m oracle.make_golden

Inputs are stored next to the outputs the unmodified reference functions produced
(bluenoise/get_noise_recent.py get_noise_v2; utils.py sample_iadb / get_scheduler_gamma),
so the fixtures stay usable where /root/reference does not exist.  ``L`` is NOT stored
(64 MiB): it is ``bndm_b200.synth.hashed_tril(seed=0)``, bit-reproducible everywhere; its
sha256 prefix is recorded in every fixture and checked by the tests.
"""
from __future__ import annotations

import hashlib
import os

import numpy as np
import torch

from bndm_b200.synth import hashed_tril
from oracle.ref_import import load_reference
from oracle.toy import ToyEps

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    get_noise_v2, _, ref_utils = load_reference()
    os.makedirs(OUT, exist_ok=True)
    L = hashed_tril(seed=0)
    sha = hashlib.sha256(L.tobytes()).hexdigest()[:16]
    Lt = torch.from_numpy(L)
    cpu = torch.device("cpu")
    rng = np.random.default_rng(20241017)

    def noise_case(name, res, bs, C, noise_type, inplace, train_or_test="train", seed=7):
        x = rng.standard_normal((bs, C, res, res)).astype(np.float32)
        gamma = rng.random(bs).astype(np.float32)
        gamma[0] = 1.0                       # gamma(T)=1: pure white
        if bs > 1:
            gamma[-1] = 0.0                  # gamma(0)=0: pure blue
        torch.manual_seed(seed)
        out, bn, wn = get_noise_v2(cpu, torch.from_numpy(x.copy()), Lt, torch.from_numpy(gamma), None,
                                   noise_type, train_or_test, inplace)
        draw = None
        if not inplace:                      # replay the draw the reference made
            torch.manual_seed(seed)
            shape = {32: (bs, C, 64, 64), 64: (bs, C, 64, 64), 128: (bs * 4, C, 64, 64)}[res]
            if noise_type == "gaussian":
                shape = x.shape
            draw = torch.randn(*shape).numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x, gamma=gamma,
                            out=out.numpy(), bn=bn.numpy(), wn=wn.numpy(),
                            draw=np.zeros(0, np.float32) if draw is None else draw,
                            noise_type=noise_type, inplace=inplace, train_or_test=train_or_test,
                            seed=seed, L_sha=sha)
        print("wrote", name, out.shape)

    noise_case("noise64_bn_b4c3_cfg1", 64, 4, 3, "gaussianBN", True)
    noise_case("noise64_gbn_b2c4", 64, 2, 4, "GBN", True)
    noise_case("noise64_rn_b2c3_draw", 64, 2, 3, "gaussianRN", False)
    noise_case("noise32_bn_b2c4", 32, 2, 4, "gaussianBN", True)
    noise_case("noise32_bn_b1c4_draw", 32, 1, 4, "gaussianBN", False)
    noise_case("noise128_bn_b2c3", 128, 2, 3, "gaussianBN", True)
    noise_case("noise128_bn_b1c3_draw", 128, 1, 3, "gaussianBN", False)
    noise_case("noise128_gauss_test_b2c3", 128, 2, 3, "gaussian", True, "test")

    # schedules: gamma / alpha tables over t = 0..T
    sched = {}
    for T in (250, 100, 1000):
        x = torch.arange(0, T + 1).float()
        sched[f"alpha_linear_T{T}"] = ref_utils.get_scheduler(x, "linear", T).numpy()
        for kind, p in (("sigmoid", (1000.0, 0.0, 3.0)), ("sigmoid", (0.2, 0.0, 3.0)),
                        ("cosine", (1.0, 0.2, 1.0)), ("linear", (1.0, 0.0, 3.0))):
            sched[f"gamma_{kind}_tau{p[0]}_s{p[1]}_e{p[2]}_T{T}"] = \
                ref_utils.get_scheduler_gamma(x, kind, p, T).numpy()
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **sched)
    print("wrote schedules", len(sched))

    # utils.sample_iadb with the toy model
    for name, nt, oc, C, T, params in (("sampler_bn_oc6", "gaussianBN", 6, 3, 250, (1000.0, 0.0, 3.0)),
                                       ("sampler_bn_oc6_tau02", "gaussianBN", 6, 3, 100, (0.2, 0.0, 3.0)),
                                       ("sampler_gauss_oc3", "gaussian", 3, 3, 250, (1000.0, 0.0, 3.0)),
                                       ("sampler_gbn_oc3_T1000", "GBN", 3, 3, 1000, (1000.0, 0.0, 3.0))):
        x0 = rng.standard_normal((3, C, 16, 16)).astype(np.float32)
        x, x_all, _ = ref_utils.sample_iadb(ToyEps(oc), torch.from_numpy(x0.copy()), T, "sigmoid", params,
                                            oc, nt, "test")
        snaps = torch.stack(x_all).numpy()
        keep = sorted(set([0, 1, len(snaps) // 2, len(snaps) - 1]))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), x0=x0, x=x.numpy(), n_snaps=len(snaps),
                            snap_idx=np.array(keep), snaps=snaps[keep], noise_type=nt, out_channel=oc,
                            nb_step=T, scheduler_params=np.array(params, np.float64))
        print("wrote", name, len(snaps))


def tiny_unet():
    """The UNet of the end-to-end ladder fixture: same block types as the reference's configs (iadb_bn.py:205-228) at
    two levels, in 3 / out 6, ~6 M parameters, deterministic init (torch.manual_seed(0), CPU)."""
    from bndm_b200.unet import UNet2DModel
    torch.manual_seed(0)
    return UNet2DModel(block_out_channels=(128, 128), in_channels=3, out_channels=6, layers_per_block=1,
                       down_block_types=("DownBlock2D", "AttnDownBlock2D"), up_block_types=("AttnUpBlock2D", "UpBlock2D")).eval()


def state_sha(model):
    h = hashlib.sha256()
    for k, v in sorted(model.state_dict().items()):
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()[:16]


def ladder_e2e():
    """Parity ladder step 4: the UNMODIFIED reference end to end on the CPU -- get_noise_v2 (inplace, per-sample
    gamma) -> utils.sample_iadb with a real (small) UNet, fp32 -- stored with its inputs; the GPU test replays it through
    the product (K1, the stock and the fused UNet, K2, eager and graphed)."""
    get_noise_v2, _, ref_utils = load_reference()
    L = hashed_tril(seed=0)
    sha = hashlib.sha256(L.tobytes()).hexdigest()[:16]
    rs = np.random.RandomState(4242)
    white = rs.randn(2, 3, 64, 64).astype(np.float32)
    gamma = np.array([1.0, 0.55], np.float32)
    model = tiny_unet()
    T, params = 6, (1000.0, 0.0, 3.0)
    with torch.no_grad():
        x0, bn, wn = get_noise_v2(torch.device("cpu"), torch.from_numpy(white.copy()), torch.from_numpy(L), torch.from_numpy(gamma),
                                  None, "gaussianBN", "test", True)
        x, x_all, _ = ref_utils.sample_iadb(model, x0, T, "sigmoid", params, 6, "gaussianBN", "test")
        d_first = model(x0, torch.full((2,), 1.0), return_dict=False)[0]
    np.savez_compressed(os.path.join(OUT, "ladder_e2e_tiny_unet.npz"), white=white, gamma=gamma, x0=x0.numpy(), x=x.numpy(),
                        snaps=torch.stack(x_all).numpy(), d_first=d_first.numpy(), nb_step=T,
                        scheduler_params=np.array(params, np.float64), L_sha=sha, model_sha=state_sha(model))
    print("wrote ladder_e2e_tiny_unet", x.shape, "model", state_sha(model))


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "ladder":
        ladder_e2e()
    else:
        main()
        ladder_e2e()
