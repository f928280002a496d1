"""DDIM sampling -- the reference's inline loop (ddim_diffusers.py:672-683, gradio_bndm.py:101-108)
as a function, with the scheduler arithmetic fused into one kernel (K3, csrc/steps.cu).

The reference delegates the arithmetic to diffusers' ``DDIMScheduler`` (constructed at
ddim_diffusers.py:499-503 with num_train_timesteps=1000, beta_schedule='linear',
prediction_type='epsilon', everything else default).  diffusers is a third-party,
unpinned dependency that is not part of the reference tree, so this module restates the
published DDIM update (Song et al. 2021) with those defaults: clip_sample=True (range 1),
set_alpha_to_one=True, timestep_spacing='leading', steps_offset=0.  PARITY UNPINNED (see
DESIGN.md); checked against oracle/sampler.py.

``sample_ddim`` does not exist in the reference; BASELINE.json names it for the inline
loop.  ``noise_fn`` lets eta > 0 steps take their variance noise from ``get_noise_v2``
(BASELINE config 3, "DDIM with time-varying noise").
"""
from __future__ import annotations

import weakref

import numpy as np
import torch

from . import _lib


class DDIMScheduler:
    """The subset of diffusers.DDIMScheduler the reference's sampling loop touches:
    ``set_timesteps``, ``timesteps``, ``step(...).prev_sample``."""

    class _Out:
        __slots__ = ("prev_sample",)

        def __init__(self, prev_sample):
            self.prev_sample = prev_sample

    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear",
                 prediction_type="epsilon", clip_sample=True):
        if beta_schedule != "linear" or prediction_type != "epsilon":
            raise NotImplementedError("only the configuration the reference uses (ddim_diffusers.py:499-503)")
        self.num_train_timesteps = num_train_timesteps
        self.beta_start, self.beta_end = float(beta_start), float(beta_end)
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.clip_sample = clip_sample
        self.num_inference_steps = None
        self.timesteps = None
        self._tables = {}

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps exceeds num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)
        if device is not None:
            self.timesteps = self.timesteps.to(device)
        self._tables.clear()

    def coefficient_table(self, eta=0.0):
        """(n,8) fp32 CPU rows {sqrt(abar_t), sqrt(1-abar_t), sqrt(abar_prev),
        sqrt(1-abar_prev-sigma^2), sigma, next timestep, 0, 0}; fp32 0-dim tensor arithmetic
        as in the scheduler."""
        ts = [int(t) for t in self.timesteps]
        ratio = self.num_train_timesteps // self.num_inference_steps
        rows = []
        for i, t in enumerate(ts):
            prev_t = t - ratio
            a_t = self.alphas_cumprod[t]
            a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
            b_t, b_prev = 1 - a_t, 1 - a_prev
            variance = (b_prev / b_t) * (1 - a_t / a_prev)
            sigma = eta * variance ** 0.5
            nxt = float(ts[i + 1]) if i + 1 < len(ts) else 0.0
            rows.append(torch.stack([a_t ** 0.5, b_t ** 0.5, a_prev ** 0.5, (1 - a_prev - sigma ** 2) ** 0.5,
                                     torch.as_tensor(sigma, dtype=torch.float32), torch.tensor(nxt),
                                     torch.tensor(0.0), torch.tensor(0.0)]))
        return torch.stack(rows).float().contiguous()

    def _table_on(self, device, eta):
        key = (str(device), float(eta))
        if key not in self._tables:
            self._tables[key] = self.coefficient_table(eta).to(device)
        return self._tables[key]

    def step(self, model_output, timestep, sample, eta: float = 0.0, variance_noise=None, generator=None):
        """prev_sample of one step; same positional surface as diffusers (model_output, timestep, sample)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        t = int(timestep)
        idx = (self.num_train_timesteps // self.num_inference_steps)
        row = self.num_inference_steps - 1 - t // idx
        if row < 0 or row >= self.num_inference_steps or int(self.timesteps[row]) != t:
            raise ValueError(f"timestep {t} is not on the inference grid")
        x = _lib.require_cuda_f32(sample, "sample")
        eps = _lib.require_cuda_f32(model_output, "model_output")
        if eta > 0 and variance_noise is None:
            variance_noise = torch.randn(eps.shape, generator=generator, device=eps.device, dtype=eps.dtype)
        noise = _lib.require_cuda_f32(variance_noise, "variance_noise") if eta > 0 else None
        table = self._table_on(x.device, eta)
        out = torch.empty_like(x)
        ddim_step_raw(out, x, eps, noise, table[row:row + 1], None, None, self.clip_sample)
        return DDIMScheduler._Out(out)


def ddim_step_raw(out, x, eps, noise, coef_rows, state, t_next_out, clip=True):
    # K3 indexes flat NCHW memory: a channels-last model output (the fused UNet's native layout) or any
    # other strided tensor is made contiguous first (a no-op for ordinary outputs)
    eps = _lib.require_cuda_f32(eps, "model output")
    if noise is not None:
        noise = _lib.require_cuda_f32(noise, "variance noise")
    with torch.cuda.device(x.device):
        rc = _lib.load().bndm_ddim_step_f32(_lib.ptr(out), _lib.ptr(x), _lib.ptr(eps), _lib.ptr(noise),
                                            _lib.ptr(coef_rows), _lib.ptr(state), _lib.ptr(t_next_out), x.shape[0],
                                            1 if clip else 0, x.numel(), _lib.current_stream(x.device))
    _lib.check(rc, "bndm_ddim_step_f32")
    return out


_graph_cache: "dict[tuple, dict]" = {}      # captured [UNet, K3] steps, see sample_ddim


@torch.no_grad()
def sample_ddim(model, x, num_inference_steps, eta=0.0, noise_fn=None, scheduler=None, use_graph=False,
                snapshots=False):
    """for t in scheduler.timesteps: eps = model(x, t).sample; x = scheduler.step(eps, t, x).prev_sample
    (ddim_diffusers.py:674-683).  ``noise_fn(i, t, x) -> variance noise`` when eta > 0.
    With ``use_graph`` (eta == 0) the [UNet, K3] pair is captured once per (model, shape, schedule) and kept:
    repeated calls replay it instead of re-capturing."""
    scheduler = scheduler or DDIMScheduler()
    scheduler.set_timesteps(num_inference_steps)
    x_in = _lib.require_cuda_f32(x, "x")
    B = x_in.shape[0]
    ts = [int(t) for t in scheduler.timesteps]

    uniform = getattr(model, "supports_uniform_timestep", False)     # one timestep for the whole batch (ddim_diffusers.py:679)

    def call(xx, tt):
        out = model(xx, tt, uniform_timestep=True) if uniform else model(xx, tt)
        return out.sample if hasattr(out, "sample") else out[0]

    seqs = []
    graph = None
    graphable = bool(use_graph) and eta == 0
    key = (id(model), tuple(x_in.shape), str(x_in.device), num_inference_steps, scheduler.num_train_timesteps,
           bool(scheduler.clip_sample), getattr(scheduler, "beta_start", None), getattr(scheduler, "beta_end", None))
    entry = _graph_cache.get(key) if graphable else None
    if entry is not None and entry["model_ref"]() is not model:       # id() reuse after garbage collection
        entry = None
    if entry is not None:
        x, table, state, t_vec, graph = entry["x"], entry["table"], entry["state"], entry["t_vec"], entry["graph"]
        x.copy_(x_in)
        state.copy_(entry["state0"])
        t_vec.fill_(float(ts[0]))
    else:
        x = x_in.clone()
        table = scheduler.coefficient_table(eta).to(x.device)
        # {block tickets of the run, number of table rows}: K3 never reads past the table (bndm_b200.h)
        state0 = torch.tensor([0, len(ts)], dtype=torch.int32, device=x.device)
        state = state0.clone()
        t_vec = torch.full((B,), float(ts[0]), dtype=torch.float32, device=x.device)
        if graphable:
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            keep = x.clone()
            with torch.cuda.stream(side):
                for _ in range(2):
                    ddim_step_raw(x, x, call(x, t_vec), None, table, state, t_vec, scheduler.clip_sample)
            torch.cuda.current_stream(x.device).wait_stream(side)
            x.copy_(keep); state.copy_(state0); t_vec.fill_(float(ts[0]))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                ddim_step_raw(x, x, call(x, t_vec), None, table, state, t_vec, scheduler.clip_sample)
            x.copy_(keep); state.copy_(state0); t_vec.fill_(float(ts[0]))
            try:
                ref = weakref.ref(model)
            except TypeError:
                ref = (lambda m: (lambda: m))(model)
            if len(_graph_cache) >= 4:
                _graph_cache.pop(next(iter(_graph_cache)))
            _graph_cache[key] = {"x": x, "table": table, "state": state, "state0": state0, "t_vec": t_vec, "graph": graph,
                                 "model_ref": ref}

    for i, t in enumerate(ts):
        if graph is not None:
            graph.replay()
        else:
            eps = _lib.require_cuda_f32(call(x, t_vec), "model output")
            vn = None
            if eta > 0:
                vn = noise_fn(i, t, x) if noise_fn is not None else torch.randn_like(eps)
                vn = _lib.require_cuda_f32(vn, "variance noise")
            ddim_step_raw(x, x, eps, vn, table, state, t_vec, scheduler.clip_sample)
        if snapshots and t % 100 == 0:                 # ddim_diffusers.py:682-683
            seqs.append(x[0:1].clone())
    if graph is not None:
        x = x.clone()                                  # the static buffer belongs to the cached graph
    return (x, seqs) if snapshots else x
