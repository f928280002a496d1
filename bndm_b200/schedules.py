"""alpha / gamma schedules and the per-step coefficient tables the step kernels read.

``get_scheduler`` / ``get_scheduler_gamma`` keep the reference's call surface
(iadb_bn.py:90-201 with the module-global ``opt``; utils.py:94-174 with explicit
``nb_steps``).  They are evaluated ONCE per sampling run on the host in torch fp32 -- the
same expressions in the same order as the reference's CPU path, because the differences
gamma(t+1)-gamma(t) are dominated by fp32 cancellation (for tau=1000 the step is the
constant 0.0039736, not 1/250) -- and uploaded as a (T,4) table.  The reference instead
re-evaluates ~20 tiny kernels plus one host->device copy per step (iadb_bn.py:306-316).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

_CLIP_MIN = 1e-9


def _curve_fraction(shape_fn, x, nb_steps, start_value, end_value):
    """1 - clamp((f(e) - f(s + t(e-s))) / (f(e) - f(s)), 1e-9, 1), t = x/nb_steps; all fp32
    tensors shaped like x, evaluated in the reference's order (iadb_bn.py:167-178)."""
    lo = torch.ones_like(x) * start_value
    hi = torch.ones_like(x) * end_value
    f_lo, f_hi = shape_fn(lo), shape_fn(hi)
    here = shape_fn((x / nb_steps) * (hi - lo) + lo)
    return 1 - torch.clamp((f_hi - here) / (f_hi - f_lo), _CLIP_MIN, 1)


def get_scheduler(x, scheduler, nb_steps, scheduler_param=1000.0):
    """alpha(t).  'linear' (utils.py:110, iadb_bn.py:106); 'sigmoid' / 'cosine' exist only in
    iadb_bn.py (:109-138) where ``scheduler_param`` is ``opt.scheduler_param``."""
    scheduler = scheduler.lower()
    if scheduler == "linear":
        return x / nb_steps
    if scheduler == "sigmoid":
        return _curve_fraction(lambda v: F.sigmoid(v / 0.9), x, nb_steps, scheduler_param, 3)
    if scheduler == "cosine":
        return _curve_fraction(lambda v: torch.cos(v * math.pi / 2) ** (2 * scheduler_param), x, nb_steps, 0.2, 1)
    raise NotImplementedError


def get_scheduler_gamma(x, scheduler, scheduler_params, nb_steps):
    """gamma(t) = white fraction; scheduler_params = (tau, start, end) (utils.py:120-174)."""
    tau, s, e = scheduler_params[0], scheduler_params[1], scheduler_params[2]
    scheduler = scheduler.lower()
    if scheduler == "linear":
        return x / nb_steps
    if scheduler == "sigmoid":
        return _curve_fraction(lambda v: F.sigmoid(v / tau), x, nb_steps, s, e)
    if scheduler == "cosine":
        return _curve_fraction(lambda v: torch.pow(torch.cos(v * math.pi / 2), 2 * tau), x, nb_steps, s, e)
    raise NotImplementedError


def iadb_table(nb_step, scheduler_alpha="linear", scheduler_gamma="sigmoid", scheduler_params=(1000.0, 0.0, 3.0),
               alpha_param=1000.0, batch=1, device="cpu", schedule_nb_steps=None):
    """(T,B,4) fp32 CPU table; row [r, b] <-> loop step t = T-1-r, sample b (iadb_bn.py:304-316):
    {alpha(t+1)-alpha(t), gamma(t+1)-gamma(t), alpha(t) [= the NEXT step's UNet timestep], 0}
    and the first UNet timestep alpha(T).

    Every step is evaluated on a (B,)-shaped int64 -> float tensor exactly like the reference
    (``tt = torch.randint(t, t + 1, (B,))``): torch's CPU kernels send short tensors and loop
    tails through scalar libm and full vectors through SLEEF, which differ in the last ulp, so
    the coefficient a sample gets depends on B and on its position in the batch.  Evaluating the
    whole schedule as one (T,) tensor would be faster and is NOT bit-identical.

    ``device``: where the schedule expressions are evaluated.  The default 'cpu' pins the table to the reference's
    CPU path (the golden vectors); pass the sampling device to reproduce a reference that itself runs on the GPU
    (iadb_bn.py:306 moves ``tt`` to ``device`` first, and CPU / CUDA sigmoid differ in the last ulp, which for
    tau = 1000 moves a per-step dgamma by up to a percent).  ``schedule_nb_steps``: the divisor of the schedules when
    it is not the loop length -- iadb_bn.py's versions divide by the global ``opt.nb_steps`` (:107, :165), whatever
    ``nb_step`` the loop was given.  Returns (table, first_t) with first_t the (B,) vector alpha(T) of the first step."""
    n_div = nb_step if schedule_nb_steps is None else schedule_nb_steps
    rows = []
    first_t = None
    for t in reversed(range(nb_step)):
        tt = torch.full((batch,), t, dtype=torch.int64).to(device)
        a_start = get_scheduler((tt + 1).float(), scheduler_alpha, n_div, alpha_param)
        a_end = get_scheduler(tt.float(), scheduler_alpha, n_div, alpha_param)
        g_start = get_scheduler_gamma((tt + 1).float(), scheduler_gamma, scheduler_params, n_div)
        g_end = get_scheduler_gamma(tt.float(), scheduler_gamma, scheduler_params, n_div)
        if first_t is None:
            first_t = a_start.float().cpu().clone()
        rows.append(torch.stack([a_start - a_end, g_start - g_end, a_end, torch.zeros_like(a_end)], dim=1))
    return torch.stack(rows).float().cpu().contiguous(), first_t


def latent_table(num_inference_steps, batch=1):
    """IADBScheduler.step coefficients (latent_iadb_bn_diffusers.py:99-103): python floats
    (t+1)/N - t/N for alpha and gamma alike, cast to fp32 when torch multiplies the tensor;
    UNet timestep alpha = (t+1)/N (:525).  Shape (N,B,4), identical across the batch."""
    N = num_inference_steps
    rows = []
    for t in reversed(range(N)):
        d = (t + 1) / N - t / N
        rows.append([d, d, t / N, 0.0])
    table = torch.tensor(rows, dtype=torch.float64).float()
    return table[:, None, :].expand(N, batch, 4).contiguous(), torch.full((batch,), float(torch.tensor(1.0 * N / N)))
