"""CPU oracle for the correlated-noise generator (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates ``get_noise_v2`` of xchhuang/bndm, bluenoise/get_noise_recent.py:23-196, twice:

``get_noise_np``     explicit index formulas in numpy (float64 accumulation by default) --
                     the independent truth the CUDA kernels are compared with.
``get_noise_torch``  the same maths with the reference's *cost profile* on torch-CPU
                     (a 2-D @ 3-D ``torch.matmul`` that ATen lowers to expand + bmm,
                     get_noise_recent.py:88,113,146) -- the timed CPU baseline of bench.py,
                     and a second parity witness with the reference's fp32 rounding.

Vocabulary: ``gamma`` is the reference's ``alpha_t`` argument = WHITE fraction per sample
(get_noise_recent.py:91,116,160): out = bn*(1-gamma) + wn*gamma.
"""
from __future__ import annotations

import numpy as np

TILE = 64            # blue-noise tile edge (cov_mat_L is (TILE*TILE)^2; iadb_bn.py:83)
NPIX = TILE * TILE

BLUE_TYPES = ("gaussianBN", "gaussianRN", "GBN")


# ----------------------------------------------------------------------------- helpers
def _lerp_f32(bn: np.ndarray, wn: np.ndarray, gamma: np.ndarray) -> np.ndarray:
    """fp32, reference association: (bn*(1-g)) + (wn*g)   (get_noise_recent.py:116)."""
    g = np.asarray(gamma, dtype=np.float32).reshape(-1, 1, 1, 1)
    one_minus = (np.float32(1.0) - g).astype(np.float32)
    return (bn * one_minus).astype(np.float32) + (wn * g).astype(np.float32)


def _apply_L(L: np.ndarray, cols: np.ndarray, acc) -> np.ndarray:
    """cols: (N, NPIX) fp32 -> (N, NPIX) fp32 with out[n, p] = sum_q L[p, q] cols[n, q]."""
    return (cols.astype(acc) @ L.astype(acc).T).astype(np.float32)


def quadrant_of_tile_in(k: int):
    """Input side of the 128^2 branch (get_noise_recent.py:131-132): tile k of the dim-0
    concatenation is x[:, :, r0:r0+64, c0:c0+64] with (r0, c0) below."""
    return ((k >> 1) * TILE, (k & 1) * TILE)


def quadrant_of_tile_out(k: int):
    """Output side (noise_padding, get_noise_recent.py:10-14): dim -2 is concatenated
    first, so tile 1 lands BELOW tile 0 and tile 2 to the right."""
    return ((k & 1) * TILE, (k >> 1) * TILE)


def split_128(x: np.ndarray) -> np.ndarray:
    """(bs,C,128,128) -> (4*bs, C, 64, 64), row index n = k*bs + b (cat on dim 0)."""
    bs = x.shape[0]
    out = np.empty((4 * bs,) + x.shape[1:2] + (TILE, TILE), dtype=x.dtype)
    for k in range(4):
        r0, c0 = quadrant_of_tile_in(k)
        out[k * bs:(k + 1) * bs] = x[:, :, r0:r0 + TILE, c0:c0 + TILE]
    return out


def place_128(tiles: np.ndarray) -> np.ndarray:
    """(bs,4,C,64,64) -> (bs,C,128,128) with the noise_padding placement."""
    bs, _, C = tiles.shape[:3]
    out = np.empty((bs, C, 2 * TILE, 2 * TILE), dtype=tiles.dtype)
    for k in range(4):
        r0, c0 = quadrant_of_tile_out(k)
        out[:, :, r0:r0 + TILE, c0:c0 + TILE] = tiles[:, k]
    return out


def scramble_white_128(z_small: np.ndarray) -> np.ndarray:
    """The 128^2 branch's ``noise_wn`` before placement (get_noise_recent.py:143-144):
    the (n, pixel, channel)-ordered copy of the white tiles re-read as (n, channel, pixel).
    wn[n, c', q] = z[n, f % C, f // C] with f = c'*NPIX + q."""
    N, C = z_small.shape[:2]
    flat = z_small.reshape(N, C, NPIX)
    f = np.arange(C * NPIX)
    wn = flat[:, f % C, f // C]                      # (N, C*NPIX) in f order
    return wn.reshape(N, C, TILE, TILE)


# ----------------------------------------------------------------------------- oracle
def get_noise_np(x, L, gamma, noise_type="gaussian", train_or_test="train", inplace=False,
                 draw=None, acc=np.float64):
    """numpy restatement of get_noise_v2 (get_noise_recent.py:23-196).

    ``draw`` stands for the tensor ``torch.randn*`` would have produced when
    ``inplace=False``: shape of x (64^2, and 128^2 'gaussian'), of the 2x2-tiled x (32^2
    branch draws AFTER tiling, :78-83) or (4*bs, C, 64, 64) (128^2 blue branch, :138).
    Returns (noise, noise_bn, noise_wn) as fp32 arrays.
    """
    x = np.asarray(x, dtype=np.float32)
    res = x.shape[-1]
    bs, C = x.shape[:2]
    if not inplace and draw is None:
        raise ValueError("inplace=False needs the white draw")

    if noise_type == "gaussian":                                   # :31-67
        if res == 64:
            z = x if inplace else np.asarray(draw, np.float32)
            return z, z, z
        if res == 128:
            z = x if inplace else np.asarray(draw, np.float32)
            if train_or_test == "test":                            # :50-56 (uses x, not z)
                tiles = scramble_white_128(split_128(x)).reshape(bs, 4, C, TILE, TILE)
                z = place_128(tiles)
            return z, z, z
        raise NotImplementedError

    if noise_type == "uniform":                                    # :69-71 then unbound at :196
        raise NotImplementedError("'uniform' never returns in the reference (unbound noise_bn)")

    if noise_type not in BLUE_TYPES:
        raise NotImplementedError

    L = np.asarray(L, dtype=np.float32)
    assert L.shape == (NPIX, NPIX)
    lerp = noise_type in ("gaussianBN", "gaussianRN")

    if res == 32:                                                  # :77-99
        z = np.tile(x, (1, 1, 2, 2)) if inplace else np.asarray(draw, np.float32)
        assert z.shape == (bs, C, TILE, TILE)
        bn = _apply_L(L, z.reshape(bs * C, NPIX), acc).reshape(bs, C, TILE, TILE)
        out = _lerp_f32(bn, z, gamma) if lerp else bn
        crop = (slice(None), slice(None), slice(0, 32), slice(0, 32))
        return out[crop], bn[crop], z[crop]

    if res == 64:                                                  # :103-120
        z = x if inplace else np.asarray(draw, np.float32)
        bn = _apply_L(L, z.reshape(bs * C, NPIX), acc).reshape(bs, C, TILE, TILE)
        out = _lerp_f32(bn, z, gamma) if lerp else bn
        return out, bn, z

    if res == 128:                                                 # :126-164
        z_small = split_128(x) if inplace else np.asarray(draw, np.float32)
        assert z_small.shape == (4 * bs, C, TILE, TILE)
        bn_small = _apply_L(L, z_small.reshape(4 * bs * C, NPIX), acc)
        # (4bs, C, ...) re-viewed as (bs, 4, C, ...): n = b'*4 + k'  (:146)
        bn = place_128(bn_small.reshape(bs, 4, C, TILE, TILE))
        wn = place_128(scramble_white_128(z_small).reshape(bs, 4, C, TILE, TILE))
        out = _lerp_f32(bn, wn, gamma) if lerp else bn
        return out, bn, wn

    raise NotImplementedError


# ------------------------------------------------------------------ timed CPU baseline
def get_noise_torch(device, x, cov_mat_L, alpha_t, time_step=None, noise_type="gaussian",
                    train_or_test="train", inplace=False):
    """torch restatement with the reference's op profile (RNG draws included).

    Same call signature as the reference.  The contraction is issued exactly as the
    reference issues it -- ``torch.matmul(L[4096,4096], z[N,4096,C])`` -- because that
    lowering (expand + bmm, L re-read once per batch entry) *is* the CPU baseline's cost.
    """
    import torch

    if noise_type == "gaussian":
        if x.shape[-1] not in (64, 128):
            raise NotImplementedError
        z = x if inplace else torch.randn_like(x)
        if x.shape[-1] == 128 and train_or_test == "test":
            z = torch.from_numpy(get_noise_np(x.cpu().numpy(), None, None, "gaussian", "test", True)[0]).to(x.device)
        return z, z, z
    if noise_type not in BLUE_TYPES:
        raise NotImplementedError
    res = x.shape[-1]
    bs, C = x.shape[:2]
    lerp = noise_type != "GBN"

    def contract(z4):                       # (N, C, 64, 64) -> L z, NCHW
        n = z4.shape[0]
        zt = z4.reshape(n, C, NPIX).transpose(1, 2)              # (n, NPIX, C) strided view
        return torch.matmul(cov_mat_L, zt).transpose(1, 2).contiguous().reshape(n, C, TILE, TILE)

    def blend(bn, wn):
        if not lerp:
            return bn
        g = alpha_t.reshape(-1, 1, 1, 1)
        return bn * (1 - g) + wn * g

    if res == 32:
        xt = x.repeat(1, 1, 2, 2)
        z = xt if inplace else torch.randn_like(xt)
        wn = z.clone()
        bn = contract(z)
        out = blend(bn, wn)
        return out[..., :32, :32], bn[..., :32, :32], wn[..., :32, :32]
    if res == 64:
        z = x if inplace else torch.randn_like(x)
        wn = z.clone()
        bn = contract(z)
        return blend(bn, wn), bn, wn
    if res == 128:
        if inplace:
            z_small = torch.cat([x[:, :, r:r + TILE, c:c + TILE]
                                 for r, c in map(quadrant_of_tile_in, range(4))], dim=0)
        else:
            z_small = torch.randn(bs * 4, C, TILE, TILE).float().to(device)
        zt = z_small.reshape(4 * bs, C, NPIX).transpose(1, 2)
        wn_t = zt.contiguous().reshape(bs, 4, C, TILE, TILE)
        bn_t = torch.matmul(cov_mat_L, zt).transpose(1, 2).contiguous().reshape(bs, 4, C, TILE, TILE)

        def place(t):
            out = t.new_empty(bs, C, 2 * TILE, 2 * TILE)
            for k in range(4):
                r0, c0 = quadrant_of_tile_out(k)
                out[:, :, r0:r0 + TILE, c0:c0 + TILE] = t[:, k]
            return out
        bn, wn = place(bn_t), place(wn_t)
        return blend(bn, wn), bn, wn
    raise NotImplementedError
