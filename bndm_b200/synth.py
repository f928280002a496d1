"""Synthetic stand-ins for the artefacts the reference downloads (README.md:33-36).

The real ``bluenoise/cov_gaussianBN_L_res64_d3.npz`` (key ``'x'``, iadb_bn.py:83) is not
redistributable/offline, so benches and tests use:

``hashed_tril``      a bit-reproducible lower-triangular fp32 matrix built from an integer
                     hash (identical on every machine -> golden vectors stay valid);
``blue_noise_L``     Cholesky factor of a toroidal high-pass ("blue") covariance on the
                     64x64 tile -- spectrally like the paper's L;
``empirical_covariance`` / ``cholesky_L``  covariance of sample fields -> L on the device;
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


def blue_noise_sigma(fc: float = 0.35, p: float = 2.0, floor: float = 1e-3) -> np.ndarray:
    """Toroidal high-pass covariance on the 64x64 tile, float64 (4096,4096), unit diagonal:
    Sigma = circulant(IFFT2(S)), S(f) = max(floor, min(1, |f|/fc)**p)."""
    f = np.fft.fftfreq(TILE)
    fr = np.sqrt(f[:, None] ** 2 + f[None, :] ** 2)
    S = np.maximum(floor, np.minimum(1.0, fr / fc) ** p)
    c = np.real(np.fft.ifft2(S))
    c /= c[0, 0]
    yy, xx = np.divmod(np.arange(NPIX), TILE)
    dy = (yy[:, None] - yy[None, :]) % TILE
    dx = (xx[:, None] - xx[None, :]) % TILE
    return c[dy, dx]


def blue_noise_L(fc: float = 0.35, p: float = 2.0, floor: float = 1e-3) -> np.ndarray:
    """cholesky(blue_noise_sigma(...)) in float64 -> fp32.  (4096,4096) lower-triangular."""
    return np.linalg.cholesky(blue_noise_sigma(fc, p, floor)).astype(np.float32)


def red_noise_sigma(fc: float = 0.08, p: float = 2.0, floor: float = 1e-3) -> np.ndarray:
    """Toroidal LOW-pass ("red") covariance on the 64x64 tile, the counterpart of ``blue_noise_sigma`` for
    ``noise_type='gaussianRN'`` (iadb_bn.py:84-85 loads cov_gaussianRN_L_res64_d3.npz):
    S(f) = max(floor, 1 / (1 + (|f|/fc)**2)**p), unit diagonal."""
    f = np.fft.fftfreq(TILE)
    fr = np.sqrt(f[:, None] ** 2 + f[None, :] ** 2)
    S = np.maximum(floor, 1.0 / (1.0 + (fr / fc) ** 2) ** p)
    c = np.real(np.fft.ifft2(S))
    c /= c[0, 0]
    yy, xx = np.divmod(np.arange(NPIX), TILE)
    dy = (yy[:, None] - yy[None, :]) % TILE
    dx = (xx[:, None] - xx[None, :]) % TILE
    return c[dy, dx]


def red_noise_L(fc: float = 0.08, p: float = 2.0, floor: float = 1e-3) -> np.ndarray:
    """cholesky(red_noise_sigma(...)) in float64 -> fp32: a synthetic stand-in for the reference's RN factor."""
    return np.linalg.cholesky(red_noise_sigma(fc, p, floor)).astype(np.float32)


def empirical_covariance(fields):
    """(S, ...) samples of a random field (e.g. S blue-noise masks of 64x64, Gaussianised) -> (n,n) float64
    covariance of the flattened, mean-removed fields, on the samples' device (torch).  The reference ships only
    the finished factor (README.md:33); this and ``cholesky_L`` are the construction it leaves out."""
    import torch
    x = fields.reshape(fields.shape[0], -1).to(torch.float64)
    x = x - x.mean(dim=0, keepdim=True)
    return (x.T @ x) / max(x.shape[0] - 1, 1)


def cholesky_L(cov, jitter: float = 0.0):
    """Covariance (n,n) -> fp32 lower-triangular factor L with L @ L.T = cov (+ jitter * I), computed in float64
    on cov's device (cuSOLVER potrf through torch on a GPU) -- the matrix ``get_noise_v2`` consumes as
    ``cov_mat_L`` (iadb_bn.py:83-86).  Raises if cov is not positive definite (add jitter for empirical estimates
    from fewer samples than pixels)."""
    import torch
    c = cov.to(torch.float64)
    if jitter:
        c = c + jitter * torch.eye(c.shape[0], dtype=torch.float64, device=c.device)
    return torch.linalg.cholesky(c).to(torch.float32)


def white_draw(shape, seed: int = 0) -> np.ndarray:
    np.random.seed(seed)
    return np.random.randn(*shape).astype(np.float32)


def save_L_npz(path: str, L: np.ndarray) -> None:
    """Writes the reference's on-disk format: key 'x' (iadb_bn.py:83)."""
    np.savez(path, x=L)
