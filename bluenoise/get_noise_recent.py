"""Import-path shim: ``from bluenoise.get_noise_recent import get_noise_v2`` (the import the
reference's scripts use, iadb_bn.py:22, latent_iadb_bn_diffusers.py:36) resolves to the
sm_100a implementation in bndm_b200.noise."""
from bndm_b200.noise import get_noise, get_noise_v2  # noqa: F401
