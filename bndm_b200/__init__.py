"""bndm_b200 -- Blackwell (sm_100a) native blue-noise diffusion sampling hot path.

Keeps the call surface of xchhuang/bndm for the one path it accelerates:

    get_noise_v2 / get_noise          bluenoise/get_noise_recent.py:23
    sample_iadb                        iadb_bn.py:287, utils.py:180
    sample_iadb_conditional            iadb_bn.py:385
    IADBScheduler, sample_latent_iadb  latent_iadb_bn_diffusers.py:75-138, :524-534
    DDIMScheduler, sample_ddim         ddim_diffusers.py:499-505, :672-683
    get_scheduler(_gamma)              utils.py:94-174
    get_model / UNet2DModel            utils.py:7-84

All compute goes through libbndm_b200.so (hand-written CUDA behind the C ABI in
include/bndm_b200.h); there is no CPU fallback.
"""
from ._lib import BndmError, LIB_PATH  # noqa: F401
from .noise import CovMatL, get_noise, get_noise_train, get_noise_v2, prepare_L  # noqa: F401
from .sampler import (GraphedModel, IADBScheduler, IadbSampler, IadbStepper, iadb_step, sample_iadb,  # noqa: F401
                      sample_iadb_conditional, sample_latent_iadb)
from .ddim import DDIMScheduler, sample_ddim  # noqa: F401
from .schedules import get_scheduler, get_scheduler_gamma, iadb_table  # noqa: F401

__version__ = "0.1.0"
