"""CPU oracle for the samplers (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates on torch (any device, normally CPU), one torch op per reference op so that fp32
rounding is the reference's:

  iadb_update              the per-step update  iadb_bn.py:323-344 / utils.py:215-228
  sample_iadb_utils        utils.sample_iadb            utils.py:179-240
  sample_iadb_opt          iadb_bn.sample_iadb          iadb_bn.py:286-379   (module-global ``opt``)
  sample_iadb_conditional  iadb_bn.py:384-438
  iadb_scheduler_step      IADBScheduler.step           latent_iadb_bn_diffusers.py:84-122
  latent_loop              latent_iadb_bn_diffusers.py:524-534 (without the VAE decode)
  DDIMTables / ddim_step / ddim_loop
                           the diffusers DDIMScheduler arithmetic that ddim_diffusers.py:499-505,
                           :639-640, :672-683 call.  diffusers is NOT under /root/reference
                           (unpinned, README.md:47) => PARITY UNPINNED; restated from the DDIM
                           update rule (Song et al., ICLR 2021, eq. 12 / 16) with the defaults
                           those call sites select: num_train_timesteps=1000,
                           beta_schedule='linear' (1e-4 .. 0.02), prediction_type='epsilon',
                           clip_sample=True (range 1), set_alpha_to_one=True,
                           timestep_spacing='leading', steps_offset=0.
"""
from __future__ import annotations

import time
from types import SimpleNamespace

import numpy as np
import torch

from .schedules import alpha_schedule, gamma_schedule

TWO_HEAD_TYPES = ("gaussianBN", "gaussianRN")


def _col(v):
    return v.view(-1, 1, 1, 1)


def iadb_update(x, d, d_alpha, d_gamma, noise_type, out_channel):
    """x + d_alpha*d[:, :C] (+ d_gamma*d[:, C:]), left-to-right (iadb_bn.py:326,329,344)."""
    C = x.shape[1]
    if noise_type in TWO_HEAD_TYPES:
        if out_channel == C:
            return x + _col(d_alpha) * d
        if out_channel == 2 * C:
            return x + _col(d_alpha) * d[:, :C, :, :] + _col(d_gamma) * d[:, C:, :, :]
        raise NotImplementedError
    if noise_type in ("gaussian", "GBN"):
        return x + _col(d_alpha) * d
    raise NotImplementedError


def _coefficients(t, batch, device, nb_step, scheduler_alpha, scheduler_gamma, scheduler_params,
                  alpha_param=1000.0):
    # iadb_bn.py:306-316: an int64 (B,) tensor of t, then four schedule evaluations
    tt = torch.full((batch,), t, dtype=torch.int64).to(device)
    a_s = alpha_schedule((tt + 1).float(), scheduler_alpha, nb_step, alpha_param)
    a_e = alpha_schedule(tt.float(), scheduler_alpha, nb_step, alpha_param)
    g_s = gamma_schedule((tt + 1).float(), scheduler_gamma, scheduler_params, nb_step)
    g_e = gamma_schedule(tt.float(), scheduler_gamma, scheduler_params, nb_step)
    return a_s, a_e, g_s, g_e


@torch.no_grad()
def sample_iadb_utils(model, x0, nb_step, scheduler_gamma, scheduler_params, out_channel,
                      noise_type, train_or_test, scheduler_alpha="linear"):
    """utils.sample_iadb (utils.py:179-240): snapshot cadence 1 (100 when nb_step==1000)."""
    return _iadb_loop(model, x0, None, nb_step, scheduler_alpha, scheduler_gamma, scheduler_params,
                      out_channel, noise_type, train_or_test, log_freq=1, with_time=True)


@torch.no_grad()
def sample_iadb_opt(model, x0, nb_step, scheduler_params, opt):
    """iadb_bn.sample_iadb (iadb_bn.py:286-379); ``opt`` is the argparse namespace the
    reference reads as a module global (:107,:311-316,:323-329,:364-373)."""
    return _iadb_loop(model, x0, None, nb_step, opt.scheduler_alpha, opt.scheduler_gamma,
                      scheduler_params, opt.out_channel, opt.noise_type, opt.train_or_test,
                      log_freq=25, with_time=True, alpha_param=getattr(opt, "scheduler_param", 1000.0),
                      schedule_nb_steps=getattr(opt, "nb_steps", nb_step))


@torch.no_grad()
def sample_iadb_conditional(model, x0, x_c, nb_step, scheduler_params, opt):
    """iadb_bn.sample_iadb_conditional (:384-438): UNet sees cat([x, x_c], 1); returns
    (x, x_all) in test mode -- no timing element (:436-437)."""
    return _iadb_loop(model, x0, x_c, nb_step, opt.scheduler_alpha, opt.scheduler_gamma,
                      scheduler_params, opt.out_channel, opt.noise_type, opt.train_or_test,
                      log_freq=25, with_time=False, alpha_param=getattr(opt, "scheduler_param", 1000.0),
                      schedule_nb_steps=getattr(opt, "nb_steps", nb_step))


def _iadb_loop(model, x0, x_c, nb_step, scheduler_alpha, scheduler_gamma, scheduler_params,
               out_channel, noise_type, train_or_test, log_freq, with_time, alpha_param=1000.0, schedule_nb_steps=None):
    # iadb_bn.py's get_scheduler / get_scheduler_gamma divide by the GLOBAL opt.nb_steps (:107, :165), whatever loop
    # length the sampler was given; utils.py's take nb_steps as an argument (utils.py:110)
    n_div = nb_step if schedule_nb_steps is None else schedule_nb_steps
    x = x0
    snaps, secs = [], []
    if nb_step == 1000:
        log_freq = 100
    for t in reversed(range(nb_step)):
        a_s, a_e, g_s, g_e = _coefficients(t, x0.shape[0], x0.device, n_div, scheduler_alpha,
                                           scheduler_gamma, scheduler_params, alpha_param)
        inp = x if x_c is None else torch.cat([x, x_c], 1)
        tic = time.time()
        d = model(inp, a_s, return_dict=False)[0]
        secs.append(time.time() - tic)
        x = iadb_update(x, d, a_s - a_e, g_s - g_e, noise_type, out_channel)
        if train_or_test == "test" and (t % log_freq == 0 or t == nb_step - 1):
            snaps.append(x)
    if train_or_test == "test":
        if with_time:
            return x, snaps, (np.mean(secs[1:]) if len(secs) > 1 else float("nan"))
        return x, snaps
    return x


# ------------------------------------------------------------------ latent IADB scheduler
def iadb_scheduler_step(model_output, timestep, x_alpha, num_inference_steps, noise_type, out_channels):
    """IADBScheduler.step (latent_iadb_bn_diffusers.py:84-122): python-float coefficients
    (t+1)/N - t/N, identical for alpha and gamma (:99-103); two heads when
    out_channels == 8 (:113)."""
    if num_inference_steps is None:
        raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' "
                         "after creating the scheduler")
    a, a_next = (timestep + 1) / num_inference_steps, timestep / num_inference_steps
    g, g_next = a, a_next
    d = model_output
    C = x_alpha.shape[1]
    if noise_type in TWO_HEAD_TYPES:
        if out_channels == C:
            return x_alpha + (a - a_next) * d
        if out_channels == 2 * C:
            return x_alpha + (a - a_next) * d[:, :C, :, :] + (g - g_next) * d[:, C:, :, :]
        raise NotImplementedError
    if noise_type == "gaussian":
        return x_alpha + (a - a_next) * d
    raise NotImplementedError


@torch.no_grad()
def latent_loop(model, noise, num_steps, noise_type, out_channels):
    """latent_iadb_bn_diffusers.py:524-534 minus the VAE decode at t==0 (weights offline)."""
    x = noise
    for t in reversed(range(num_steps)):
        alpha = (t + 1) / num_steps
        d = model(x, torch.tensor(alpha, device=x.device), return_dict=False)[0]
        x = iadb_scheduler_step(d, t, x, num_steps, noise_type, out_channels)
    return x


# ------------------------------------------------------------------ DDIM (parity unpinned)
class DDIMTables:
    """alpha-bar table + 'leading' timestep grid of the diffusers DDIMScheduler as the
    reference configures it (ddim_diffusers.py:499-503, :640)."""

    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02):
        self.num_train_timesteps = num_train_timesteps
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = None

    def set_timesteps(self, n):
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        self.timesteps = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64)

    def coefficients(self, t, eta=0.0):
        """(sqrt(abar_t), sqrt(1-abar_t), sqrt(abar_prev), sqrt(1-abar_prev-sigma^2), sigma)
        as fp32 0-dim tensors, each derived in fp32 like the 0-dim-tensor arithmetic of the
        scheduler."""
        t = int(t)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        variance = (b_prev / b_t) * (1 - a_t / a_prev)
        sigma = eta * variance ** 0.5
        return a_t ** 0.5, b_t ** 0.5, a_prev ** 0.5, (1 - a_prev - sigma ** 2) ** 0.5, sigma


def ddim_step(tables: DDIMTables, eps, t, x, eta=0.0, variance_noise=None, clip_sample=True):
    """prev_sample of one DDIM step (epsilon prediction)."""
    sa, sb, sap, sdir, sigma = tables.coefficients(t, eta)
    x0 = (x - sb * eps) / sa
    if clip_sample:
        x0 = x0.clamp(-1.0, 1.0)
    direction = sdir * eps
    prev = sap * x0 + direction
    if eta > 0:
        if variance_noise is None:
            variance_noise = torch.randn_like(eps)
        prev = prev + sigma * variance_noise
    return prev


@torch.no_grad()
def ddim_loop(model, x, num_inference_steps, eta=0.0, noise_fn=None, tables=None):
    """ddim_diffusers.py:672-683: for t in scheduler.timesteps: eps = model(x, t).sample;
    x = scheduler.step(eps, t, x).prev_sample.  ``noise_fn(i, t, x)`` supplies the variance
    noise when eta > 0 (BASELINE config 3 composes it from get_noise_v2)."""
    tables = tables or DDIMTables()
    tables.set_timesteps(num_inference_steps)
    for i, t in enumerate(tables.timesteps):
        out = model(x, torch.tensor(int(t), device=x.device))
        eps = out.sample if hasattr(out, "sample") else out[0]
        vn = noise_fn(i, int(t), x) if (eta > 0 and noise_fn is not None) else None
        x = ddim_step(tables, eps, int(t), x, eta, vn)
    return x


def make_opt(**kw):
    """An argparse-like namespace with the hot-path flags of iadb_bn.py:29-69."""
    base = dict(noise_type="gaussianBN", out_channel=6, scheduler_alpha="linear",
                scheduler_gamma="sigmoid", scheduler_param=1000.0, train_or_test="test",
                nb_steps=250)
    base.update(kw)
    return SimpleNamespace(**base)
