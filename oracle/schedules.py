"""CPU oracle for the alpha / gamma schedules (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, on torch-CPU fp32 and with the reference's order of operations (the differences
gamma(t+1)-gamma(t) are dominated by fp32 cancellation, SURVEY.md App. B, so the order IS
the contract):

  get_scheduler        iadb_bn.py:90-143   ('linear' :106-107, 'sigmoid' :109-125,
                                            'cosine' :127-138) and utils.py:94-116
  get_scheduler_gamma  iadb_bn.py:147-201, utils.py:120-174
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _normalised_curve(f, x, nb_steps, s, e):
    """1 - clamp((f(e) - f(t*(e-s)+s)) / (f(e) - f(s)), 1e-9, 1) with t = x/nb_steps,
    every intermediate an fp32 tensor shaped like x (iadb_bn.py:167-178, :180-194)."""
    lo = torch.ones_like(x) * s
    hi = torch.ones_like(x) * e
    f_lo, f_hi = f(lo), f(hi)
    t = x / nb_steps
    cur = f(t * (hi - lo) + lo)
    frac = (f_hi - cur) / (f_hi - f_lo)
    return 1 - torch.clamp(frac, 1e-9, 1)


def alpha_schedule(x: torch.Tensor, kind: str, nb_steps: int, scheduler_param: float = 1000.0):
    """iadb_bn.get_scheduler (:90-143).  ``scheduler_param`` is ``opt.scheduler_param``.
    utils.get_scheduler (utils.py:94-116) implements only 'linear'."""
    kind = kind.lower()
    if kind == "linear":
        return x / nb_steps
    if kind == "sigmoid":                     # start=opt.scheduler_param, end=3, tau=0.9 (:115-119)
        return _normalised_curve(lambda v: F.sigmoid(v / 0.9), x, nb_steps, scheduler_param, 3)
    if kind == "cosine":                      # start=0.2, end=1, tau=opt.scheduler_param (:128-131)
        tau = scheduler_param
        return _normalised_curve(lambda v: torch.cos(v * math.pi / 2) ** (2 * tau), x, nb_steps, 0.2, 1)
    raise NotImplementedError(kind)


def gamma_schedule(x: torch.Tensor, kind: str, scheduler_params, nb_steps: int):
    """get_scheduler_gamma (iadb_bn.py:147-201 == utils.py:120-174).
    scheduler_params = (tau, start, end)."""
    tau, s, e = scheduler_params[0], scheduler_params[1], scheduler_params[2]
    kind = kind.lower()
    if kind == "linear":
        return x / nb_steps
    if kind == "sigmoid":
        return _normalised_curve(lambda v: F.sigmoid(v / tau), x, nb_steps, s, e)
    if kind == "cosine":
        # NB the reference writes pi/2.0 and 2.0*tau for v_start but pi/2 and 2*tau elsewhere
        # (:186-190) -- identical values in fp32.
        return _normalised_curve(lambda v: torch.pow(torch.cos(v * math.pi / 2), 2 * tau), x, nb_steps, s, e)
    raise NotImplementedError(kind)
