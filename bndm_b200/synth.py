"""Synthetic stand-ins for the artefacts the reference downloads (README.md:33-36).

The real ``bluenoise/cov_gaussianBN_L_res64_d3.npz`` (key ``'x'``, iadb_bn.py:83) is not
redistributable/offline, so benches and tests use:

``hashed_tril``      a bit-reproducible lower-triangular fp32 matrix built from an integer
                     hash (identical on every machine -> golden vectors stay valid);
``blue_noise_L``     Cholesky factor of a toroidal high-pass ("blue") covariance on the
                     64x64 tile -- spectrally like the paper's L;
``white_draw``       ``np.random.seed(s); np.random.randn(...)`` exactly as iadb_bn.py:75,761.
"""
from __future__ import annotations

import numpy as np

TILE = 64
NPIX = TILE * TILE


def hashed_tril(n: int = NPIX, seed: int = 0) -> np.ndarray:
    """Lower-triangular (n,n) fp32; entries are k * 2**-17 with integer k in [-2048, 2047],
    so every value (and the construction) is exact in fp32 on any platform."""
    i = np.arange(n, dtype=np.uint64)[:, None]
    j = np.arange(n, dtype=np.uint64)[None, :]
    h = i * np.uint64(0x9E3779B1) + j * np.uint64(0x85EBCA77) + np.uint64(seed * 0x27D4EB2F + 0x165667B1)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x2C1B3C6D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(12)
    h = (h * np.uint64(0x297A2D39)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    k = (h & np.uint64(0xFFF)).astype(np.int64) - 2048
    L = (k.astype(np.float32) * np.float32(2.0 ** -17))
    return np.tril(L).astype(np.float32)


def blue_noise_L(fc: float = 0.35, p: float = 2.0, floor: float = 1e-3) -> np.ndarray:
    """cholesky(Sigma) in float64 -> fp32, Sigma = circulant(IFFT2(S)), S(f) = max(floor,
    min(1, |f|/fc)**p), unit diagonal.  (4096,4096) lower-triangular."""
    f = np.fft.fftfreq(TILE)
    fr = np.sqrt(f[:, None] ** 2 + f[None, :] ** 2)
    S = np.maximum(floor, np.minimum(1.0, fr / fc) ** p)
    c = np.real(np.fft.ifft2(S))
    c /= c[0, 0]
    yy, xx = np.divmod(np.arange(NPIX), TILE)
    dy = (yy[:, None] - yy[None, :]) % TILE
    dx = (xx[:, None] - xx[None, :]) % TILE
    sigma = c[dy, dx]
    return np.linalg.cholesky(sigma).astype(np.float32)


def white_draw(shape, seed: int = 0) -> np.ndarray:
    np.random.seed(seed)
    return np.random.randn(*shape).astype(np.float32)


def save_L_npz(path: str, L: np.ndarray) -> None:
    """Writes the reference's on-disk format: key 'x' (iadb_bn.py:83)."""
    np.savez(path, x=L)
