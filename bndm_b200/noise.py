"""Correlated (blue) noise generator -- drop-in for the reference's ``get_noise_v2``
(bluenoise/get_noise_recent.py:23-196) backed by the sm_100a kernels in csrc/.

Same call surface and return triple ``(noise, noise_bn, noise_wn)``; the white draw stays
in torch (``torch.randn_like`` / ``torch.randn`` exactly where the reference draws, so
identical seeds give identical white fields); everything after the draw -- contraction
with L, transposes, 32^2 tiling/crop, 128^2 tiling and white re-interpretation, the
white<->blue lerp -- runs in three launches of libbndm_b200.so (pack, tcgen05 GEMM,
epilogue).  CUDA tensors only.
"""
from __future__ import annotations

import weakref

import torch

from . import _lib

BLUE_TYPES = ("gaussianBN", "gaussianRN", "GBN")
_GEMM_FLAGS = {"auto": _lib.GEMM_AUTO, "tc": _lib.GEMM_TC, "simt": _lib.GEMM_SIMT, "gemv": _lib.GEMM_GEMV}


class CovMatL:
    """Device-side handle of the Cholesky factor ``cov_mat_L`` (iadb_bn.py:83-86): keeps the
    caller's tensor alive plus the kernel-side operand copies and workspace."""

    def __init__(self, cov_mat_L: torch.Tensor, max_columns: int = 192):
        L = _lib.require_cuda_f32(cov_mat_L, "cov_mat_L")
        if L.dim() != 2 or L.shape[0] != L.shape[1]:
            raise ValueError(f"cov_mat_L must be square, got {tuple(L.shape)}")
        self.tensor = L
        self.device = L.device
        lib = _lib.load()
        handle = _lib.C.c_void_p()
        with torch.cuda.device(self.device):
            rc = lib.bndm_prepare_L(_lib.ptr(L), int(L.shape[0]), int(max_columns),
                                    _lib.current_stream(self.device), _lib.C.byref(handle))
        _lib.check(rc, "bndm_prepare_L")
        self._h = handle
        self.profile_enabled = False
        self._finalizer = weakref.finalize(self, lib.bndm_free_L, handle)

    @property
    def lower_triangular(self) -> bool:
        return bool(_lib.load().bndm_L_is_lower_triangular(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(_lib.load().bndm_workspace_bytes(self._h))

    def reserve(self, max_columns: int):
        """Pre-size the workspace (needed before CUDA-graph capture of get_noise_v2)."""
        with torch.cuda.device(self.device):
            rc = _lib.load().bndm_reserve_columns(self._h, int(max_columns), _lib.current_stream(self.device))
        _lib.check(rc, "bndm_reserve_columns")

    def profile(self, on=True):
        _lib.check(_lib.load().bndm_profile_enable(self._h, 1 if on else 0), "bndm_profile_enable")
        self.profile_enabled = bool(on)

    def last_ms(self):
        """(pack, contraction, epilogue) milliseconds of the last profiled get_noise call."""
        a, b, c = _lib.C.c_float(), _lib.C.c_float(), _lib.C.c_float()
        _lib.check(_lib.load().bndm_profile_last_ms(self._h, _lib.C.byref(a), _lib.C.byref(b), _lib.C.byref(c)),
                   "bndm_profile_last_ms")
        return a.value, b.value, c.value

    def close(self):
        self._finalizer()


_handles: "dict[tuple, tuple]" = {}      # (data_ptr, device, shape) -> (_version, handle); insertion order = age
_MAX_HANDLES = 4


def prepare_L(cov_mat_L, max_columns: int = 192) -> CovMatL:
    """Returns the cached handle for this matrix.  A cached handle keeps the caller's tensor object alive (``_orig``,
    next to the contiguous fp32 tensor the kernels read), so the memory a key points at cannot be freed and handed
    to another tensor while the entry exists: a new L at a recycled address can never inherit stale operand copies.
    In-place modification (the tensor's ``_version``) drops the entry; at most 4 matrices are kept."""
    if isinstance(cov_mat_L, CovMatL):
        return cov_mat_L
    if not isinstance(cov_mat_L, torch.Tensor):
        raise TypeError("cov_mat_L must be a torch.Tensor or CovMatL")
    key = (cov_mat_L.data_ptr(), str(cov_mat_L.device), tuple(cov_mat_L.shape), tuple(cov_mat_L.stride()))
    hit = _handles.get(key)
    if hit is not None:
        if hit[0] == cov_mat_L._version:
            return hit[1]
        _handles.pop(key)                                         # modified in place since the copies were made
    while len(_handles) >= _MAX_HANDLES:
        _handles.pop(next(iter(_handles)))                        # dropped, not closed: a caller may still hold the
                                                                  # handle; its finalizer frees it with the last reference
    h = CovMatL(cov_mat_L, max_columns)
    h._orig = cov_mat_L                                           # pins the keyed memory for the lifetime of the entry
    _handles[key] = (cov_mat_L._version, h)
    return h


def _white128_reinterpret(x, lo=0, count=None):
    count = x.shape[0] if count is None else count
    out = torch.empty((count,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().bndm_white128_reinterpret_shard_f32(_lib.ptr(x), _lib.ptr(out), count, x.shape[1], x.shape[0], lo,
                                                             _lib.current_stream(x.device))
    _lib.check(rc, "bndm_white128_reinterpret_shard_f32")
    return out


def get_noise_v2(device, x, cov_mat_L, alpha_t, time_step=None, noise_type="gaussian", train_or_test="train",
                 inplace=False, *, gemm="auto", want=("noise", "bn", "wn"), shard=None):
    """Drop-in for bluenoise/get_noise_recent.py:23.

    ``alpha_t`` is the per-sample WHITE fraction gamma (B,) -- out = bn*(1-gamma) + wn*gamma
    (:91,:116,:160).  ``time_step`` is unused, as in the reference.  Extra keyword-only
    arguments: ``gemm`` in {'auto' (K1g streaming fp32 kernel for <= 16 columns, tcgen05 above), 'tc' (tcgen05
    3xTF32), 'gemv' (K1g), 'simt' (fp32 FFMA split-K witness)}; ``want`` lets callers
    that ignore noise_bn / noise_wn skip writing them (those entries are returned as None);
    ``shard=(lo, hi)``: ``x`` (and ``alpha_t``, if it has one entry per global sample) describe the WHOLE batch
    of a run that is split across GPUs and the call returns samples ``[lo, hi)`` only, bit-identical to
    the same rows of the unsharded call -- including the 128^2 ``inplace=True`` branch, which mixes
    samples across the batch (:131-146), and the draw, which is made for the global shape so that
    identical seeds give identical fields on any number of GPUs.
    """
    if x.dim() != 4:
        raise ValueError("x must be (B, C, H, W)")
    res = x.shape[-1]
    bs_g, dimension = x.shape[0], x.shape[1]
    lo, hi = (0, bs_g) if shard is None else (int(shard[0]), int(shard[1]))
    if not (0 <= lo < hi <= bs_g):
        raise ValueError(f"shard {shard} outside the batch of {bs_g}")
    bs = hi - lo

    if noise_type == "gaussian":                                   # :31-67, pass-through
        if res not in (64, 128):
            raise NotImplementedError
        noise = x if inplace else torch.randn_like(x)
        if res == 128 and train_or_test == "test":                 # :50-56 (built from x, not the draw)
            noise = _white128_reinterpret(_lib.require_cuda_f32(x, "x"), lo, bs)
        elif shard is not None:
            noise = noise[lo:hi]
        return noise, noise, noise

    if noise_type == "uniform":
        # the reference computes a uniform field then fails with an unbound `noise_bn` at :196
        raise NotImplementedError("noise_type='uniform' never returns in the reference")
    if noise_type not in BLUE_TYPES:
        raise NotImplementedError
    if res not in (32, 64, 128):
        raise NotImplementedError

    x = _lib.require_cuda_f32(x, "x")
    if x.shape[-2] != res:
        raise ValueError("x must be square")
    L = prepare_L(cov_mat_L, max_columns=bs * dimension * (4 if res == 128 else 1))
    if L.device != x.device:
        raise ValueError(f"cov_mat_L is on {L.device}, x on {x.device}")

    # ---- the white field: drawn where and how the reference draws it (for the GLOBAL batch)
    if inplace:
        z, src = x, _lib.SRC_IMAGE
    else:
        src = _lib.SRC_DRAW
        if res == 64:
            z = torch.randn_like(x)                                                    # :108
        elif res == 32:
            z = torch.randn(bs_g, dimension, 64, 64, dtype=x.dtype, device=x.device)   # :78-83 randn_like(tiled x)
        else:
            z = torch.randn(bs_g * 4, dimension, 64, 64).float().to(device)            # :138 (CPU generator)
            z = _lib.require_cuda_f32(z, "white draw")

    gamma = None
    if noise_type in ("gaussianBN", "gaussianRN"):
        gamma = _lib.require_cuda_f32(alpha_t.reshape(-1), "alpha_t")
        if shard is not None and gamma.numel() == bs_g:
            gamma = gamma[lo:hi].contiguous()
        if gamma.numel() != bs:
            raise ValueError(f"alpha_t must have {bs} entries, got {gamma.numel()}")

    shape = (bs, dimension, res, res)
    out = torch.empty(shape, dtype=torch.float32, device=x.device)
    out_bn = torch.empty_like(out) if (gamma is not None and "bn" in want) else None
    out_wn = torch.empty_like(out) if "wn" in want else None
    with torch.cuda.device(x.device):
        rc = _lib.load().bndm_get_noise_shard_f32(L._h, _lib.ptr(z), _lib.ptr(gamma), _lib.ptr(out), _lib.ptr(out_bn),
                                                  _lib.ptr(out_wn), bs, dimension, res, src | _GEMM_FLAGS[gemm], bs_g, lo,
                                                  _lib.current_stream(x.device))
    _lib.check(rc, "bndm_get_noise_f32")
    if gamma is None:          # 'GBN': noise IS noise_bn (:118)
        out_bn = out
    return out, out_bn, out_wn


get_noise = get_noise_v2   # the name BASELINE.json's north_star uses


def get_noise_train(device, x1, cov_mat_L, gamma_t, alpha, alpha_prev=None, noise_type="gaussianBN", *, draw=None,
                    want_x0=False, gemm="auto"):
    """The noise + blend front end of one IADB training step (iadb_bn.py:881-954) in one call:

        x0, bn, wn = get_noise_v2(device, x1, L, gamma_t, t, noise_type, 'train', inplace=False)
        x_alpha = alpha*x0 + (1-alpha)*x1;  tar1 = x1 - x0;  tar2 = alpha_prev*(bn - wn)

    Returns ``(x_alpha, tar1, tar2[, x0])`` (``tar2`` is None for 'GBN' or when ``alpha_prev`` is None).  The
    white field is drawn exactly where get_noise_v2 draws it (``draw`` overrides it, for tests).  The
    contraction is the same kernel; only its epilogue differs (5 element-wise passes and 3 temporaries
    less than the torch sequence)."""
    if noise_type not in BLUE_TYPES:
        raise NotImplementedError
    x1 = _lib.require_cuda_f32(x1, "x1")
    bs, C, res = x1.shape[0], x1.shape[1], x1.shape[-1]
    if res not in (32, 64, 128):
        raise NotImplementedError
    L = prepare_L(cov_mat_L, max_columns=bs * C * (4 if res == 128 else 1))
    if draw is not None:
        z = _lib.require_cuda_f32(draw, "draw")
    elif res == 64:
        z = torch.randn_like(x1)
    elif res == 32:
        z = torch.randn(bs, C, 64, 64, dtype=x1.dtype, device=x1.device)
    else:
        z = _lib.require_cuda_f32(torch.randn(bs * 4, C, 64, 64).float().to(device), "white draw")
    gamma = None
    if noise_type in ("gaussianBN", "gaussianRN"):
        gamma = _lib.require_cuda_f32(gamma_t.reshape(-1), "gamma_t")
    alpha = _lib.require_cuda_f32(alpha.reshape(-1), "alpha")
    two = gamma is not None and alpha_prev is not None
    if two:
        alpha_prev = _lib.require_cuda_f32(alpha_prev.reshape(-1), "alpha_prev")
    x_alpha, tar1 = torch.empty_like(x1), torch.empty_like(x1)
    tar2 = torch.empty_like(x1) if two else None
    x0 = torch.empty_like(x1) if want_x0 else None
    with torch.cuda.device(x1.device):
        rc = _lib.load().bndm_get_noise_train_f32(L._h, _lib.ptr(z), _lib.ptr(gamma), _lib.ptr(x1), _lib.ptr(alpha),
                                                  _lib.ptr(alpha_prev if two else None), _lib.ptr(x_alpha), _lib.ptr(tar1),
                                                  _lib.ptr(tar2), _lib.ptr(x0), bs, C, res, _lib.SRC_DRAW | _GEMM_FLAGS[gemm],
                                                  _lib.current_stream(x1.device))
    _lib.check(rc, "bndm_get_noise_train_f32")
    return (x_alpha, tar1, tar2, x0) if want_x0 else (x_alpha, tar1, tar2)
