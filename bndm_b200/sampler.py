"""IADB samplers -- drop-ins for the reference's sampling loops, with the per-step update
done by one fused kernel of libbndm_b200.so and (optionally) the whole
[UNet forward -> update] step replayed from one CUDA graph.

Reference surfaces kept:
  sample_iadb(model, x0, nb_step, scheduler_params)                        iadb_bn.py:287 (reads module ``opt``)
  sample_iadb(model, x0, nb_step, scheduler_gamma, scheduler_params,
              out_channel, noise_type, train_or_test, scheduler_alpha)     utils.py:180
  sample_iadb_conditional(model, x0, x_c, nb_step, scheduler_params)       iadb_bn.py:385
  IADBScheduler().set_timesteps(n) / .step(model_output, t, x_alpha)       latent_iadb_bn_diffusers.py:75-125
  sample_latent_iadb(...)                                                  latent_iadb_bn_diffusers.py:524-534

What changed underneath: the reference builds the step coefficients with ~20 tiny kernels
and one host->device copy per step (iadb_bn.py:306-316) and applies them with 3-5
element-wise launches (:326-344).  Here the coefficients are a device table computed once
(schedules.iadb_table), and K2 (csrc/steps.cu) applies a row, writes the next UNet
timestep and advances the device-side step counter -- so a step has no host work and one
captured graph serves all T steps.
"""
from __future__ import annotations

import time
import weakref
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .schedules import iadb_table, latent_table

TWO_HEAD = ("gaussianBN", "gaussianRN")

# Module-global options namespace, like iadb_bn.py:69 (`opt = parser.parse_args()`).
opt = SimpleNamespace(noise_type="gaussianBN", out_channel=6, scheduler_alpha="linear", scheduler_gamma="sigmoid",
                      scheduler_param=1000.0, train_or_test="test", nb_steps=250)


def _expected_out_channels(noise_type, out_channel, C):
    """Mirrors the branch structure (and NotImplementedError sites) of iadb_bn.py:323-346."""
    if noise_type in TWO_HEAD:
        if out_channel in (C, 2 * C):
            return out_channel
        raise NotImplementedError
    if noise_type in ("gaussian", "GBN"):
        return C
    raise NotImplementedError


class IadbStepper:
    """Device-resident schedule + K2 launcher for one sampling run of fixed shape."""

    def __init__(self, table_cpu: torch.Tensor, first_t, batch: int, device, expect_channels=None):
        """``first_t``: the UNet timestep of the first step -- a float, or the (B,) vector the reference forms
        (alpha_start per sample, iadb_bn.py:311).  ``expect_channels``: channel count the model output must have
        (the reference fails with a shape error on anything else, iadb_bn.py:326-344)."""
        self.device = torch.device(device)
        if table_cpu.dim() != 3 or table_cpu.shape[1] != batch or table_cpu.shape[2] != 4:
            raise ValueError(f"schedule table must be (T, {batch}, 4), got {tuple(table_cpu.shape)}")
        self.n_steps = table_cpu.shape[0]
        self.table = table_cpu.to(self.device).contiguous()
        # {block tickets of the run, number of table rows}: K2 never reads past the table (see bndm_b200.h)
        self._state0 = torch.tensor([0, self.n_steps], dtype=torch.int32, device=self.device)
        self.state = self._state0.clone()
        first = torch.as_tensor(first_t, dtype=torch.float32).reshape(-1)
        self._first_t = (first.expand(batch) if first.numel() == 1 else first).contiguous().to(self.device)
        if self._first_t.numel() != batch:
            raise ValueError(f"first_t must be a scalar or have {batch} entries")
        self.t_vec = self._first_t.clone()
        self.expect_channels = expect_channels
        self._issued = 0          # steps issued since reset (host side; replays of a captured step count too)
        self._nhwc = None         # kernel variant of this run (the grid size enters the device-side step index)

    def reset(self):
        self.state.copy_(self._state0)
        self.t_vec.copy_(self._first_t)
        self._issued = 0
        self._nhwc = None

    @property
    def overrun(self) -> bool:
        """True if a step was launched past the end of the schedule since the last reset (device-side flag)."""
        return bool(int(self.state[1].item()) & 0x40000000)

    def note_step(self, n=1):
        self._issued += n
        if self._issued > self.n_steps:
            raise RuntimeError(f"IadbStepper: step {self._issued} of a {self.n_steps}-step schedule -- call reset() "
                               f"before starting another run")

    def step_(self, x: torch.Tensor, d: torch.Tensor):
        """x <- x + dalpha*d[:, :C] (+ dgamma*d[:, C:]) in place; advances t_vec / state."""
        B, C = x.shape[0], x.shape[1]
        lib = _lib.load()
        if not (isinstance(d, torch.Tensor) and d.dim() == 4 and d.shape[0] == B and d.shape[2:] == x.shape[2:]):
            raise ValueError(f"model output {tuple(getattr(d, 'shape', ()))} does not match x {tuple(x.shape)}")
        if self.expect_channels is not None and d.shape[1] != self.expect_channels:
            raise ValueError(f"model output has {d.shape[1]} channels, the schedule expects {self.expect_channels} "
                             f"(iadb_bn.py:326-344 fails on this shape)")
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise ValueError("x must be a contiguous float32 CUDA tensor (updated in place)")
        self.note_step()
        # the channels-last UNet evaluation hands over d in NHWC memory: consumed in place
        nhwc = (d.is_cuda and d.dtype == torch.float32 and not d.is_contiguous()
                and d.is_contiguous(memory_format=torch.channels_last))
        if self._nhwc is None:
            self._nhwc = nhwc
        elif self._nhwc != nhwc:
            raise RuntimeError("model output changed memory format within a run (the step kernel variant is fixed per run)")
        if not nhwc:
            d = _lib.require_cuda_f32(d, "model output")
        fn = lib.bndm_iadb_step_sched_dnhwc_f32 if nhwc else lib.bndm_iadb_step_sched_f32
        with torch.cuda.device(self.device):
            rc = fn(_lib.ptr(x), _lib.ptr(x), _lib.ptr(d), _lib.ptr(self.table), _lib.ptr(self.state),
                    _lib.ptr(self.t_vec), B, C, x.shape[2] * x.shape[3], d.shape[1], _lib.current_stream(self.device))
        _lib.check(rc, "bndm_iadb_step_sched_f32")
        return x


def iadb_step(x, d, dalpha, dgamma=None, out=None):
    """One update with per-sample coefficient vectors (the tensors the reference forms at
    iadb_bn.py:311-316): returns (x + dalpha*d[:, :C]) + dgamma*d[:, C:]."""
    x = _lib.require_cuda_f32(x, "x")
    d = _lib.require_cuda_f32(d, "d")
    dalpha = _lib.require_cuda_f32(dalpha.reshape(-1), "dalpha")
    if dgamma is not None:
        dgamma = _lib.require_cuda_f32(dgamma.reshape(-1), "dgamma")
    if d.shape[0] != x.shape[0] or d.shape[2:] != x.shape[2:]:
        raise ValueError(f"d {tuple(d.shape)} does not match x {tuple(x.shape)}")
    if out is None:
        out = torch.empty_like(x)
    elif not (isinstance(out, torch.Tensor) and out.is_cuda and out.device == x.device and out.dtype == torch.float32
              and out.is_contiguous() and out.shape == x.shape):
        raise ValueError("out must be a contiguous float32 CUDA tensor shaped like x on x's device")
    B, C = x.shape[0], x.shape[1]
    with torch.cuda.device(x.device):
        rc = _lib.load().bndm_iadb_step_f32(_lib.ptr(out), _lib.ptr(x), _lib.ptr(d), _lib.ptr(dalpha), _lib.ptr(dgamma),
                                            B, C, x.shape[2] * x.shape[3], d.shape[1], _lib.current_stream(x.device))
    _lib.check(rc, "bndm_iadb_step_f32")
    return out


class GraphedStep:
    """One sampling step for a fixed shape, captured once and replayed per step; x, t_vec and
    the step counter are static device buffers the graph reads and writes.

    fused=True   graph = [d = model(x, t_vec); K2(x, d)]            (one replay per step)
    fused=False  graph = [d = model(x, t_vec)], K2 launched eagerly after each replay on the
                 same stream -- lets a caller bracket K2 with CUDA events (bench.py's live
                 roofline measurement); same kernels, same order, same results."""

    def __init__(self, model_call, stepper: IadbStepper, x_static: torch.Tensor, warmup: int = 2, fused: bool = True):
        self.stepper = stepper
        self.x = x_static
        self.fused = fused
        self.d = None
        self.graph = torch.cuda.CUDAGraph()
        dev = x_static.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        keep = x_static.clone()
        with torch.cuda.stream(side):
            for _ in range(warmup):                       # lazy inits (cuDNN plans, workspaces) outside capture
                stepper.reset()                           # (a 1-step schedule must not count the warm-ups as its run)
                stepper.step_(x_static, model_call(x_static, stepper.t_vec))
        torch.cuda.current_stream(dev).wait_stream(side)
        x_static.copy_(keep)
        stepper.reset()
        with torch.cuda.graph(self.graph, stream=side):
            d = model_call(x_static, stepper.t_vec)
            if fused:
                stepper.step_(x_static, d)
            else:
                self.d = d
        stepper.reset()

    def replay(self, events=None):
        if self.fused:
            self.stepper.note_step()          # the captured K2 advances the device-side step index
        self.graph.replay()
        if not self.fused:
            if events is not None:
                events[0].record()
            self.stepper.step_(self.x, self.d)
            if events is not None:
                events[1].record()


class GraphedLoop:
    """The WHOLE sampling run -- T x [d = model(x, t_vec); K2(x, d)] -- captured as one CUDA graph: one replay per run,
    no host work at all between the first UNet launch and the last update.  Possible because K2 keeps the
    step index on the device and publishes the next timestep itself.  Capture costs T eager forwards once per
    (model, shape, schedule); the graph's private pool is reused across the T iterations, so memory is one forward's."""

    def __init__(self, model_call, stepper: "IadbStepper", x_static: torch.Tensor, n_steps: int, warmup: int = 2):
        self.stepper, self.x, self.n_steps = stepper, x_static, n_steps
        self.graph = torch.cuda.CUDAGraph()
        dev = x_static.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        keep = x_static.clone()
        with torch.cuda.stream(side):
            for _ in range(warmup):                       # lazy inits (cuDNN plans, workspaces) outside capture
                stepper.reset()
                stepper.step_(x_static, model_call(x_static, stepper.t_vec))
        torch.cuda.current_stream(dev).wait_stream(side)
        x_static.copy_(keep)
        stepper.reset()
        with torch.cuda.graph(self.graph, stream=side):
            for _ in range(n_steps):
                stepper.step_(x_static, model_call(x_static, stepper.t_vec))
        stepper.reset()

    def replay(self):
        self.stepper.note_step(self.n_steps)
        self.graph.replay()


class GraphedModel:
    """``model(x, t)`` for one fixed input shape replayed from a CUDA graph (north_star: "the UNet2DModel forward ...
    wrapped with CUDA Graphs per fixed shape").  Keeps the diffusers call conventions -- ``(x, t).sample`` and
    ``(x, t, return_dict=False)[0]`` -- for loops that cannot be captured as a whole (DDIM with eta > 0 draws fresh
    variance noise through the host every step).  The returned tensor is the graph's static output buffer: consume it
    before the next call."""

    def __init__(self, model, shape, device="cuda", t_dtype=torch.float32, warmup=2, uniform_timestep=False):
        """``uniform_timestep``: the loop passes one timestep for the whole batch (see FusedUNet2D.forward)."""
        self.device = torch.device(device)
        kw = {"uniform_timestep": True} if (uniform_timestep and getattr(model, "supports_uniform_timestep", False)) else {}
        self.x = torch.zeros(tuple(shape), dtype=torch.float32, device=self.device)
        self.t = torch.zeros(shape[0], dtype=t_dtype, device=self.device)
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.no_grad():
            with torch.cuda.stream(side):
                for _ in range(warmup):                       # lazy inits (cuDNN plans, workspaces) outside capture
                    model(self.x, self.t, return_dict=False, **kw)
            torch.cuda.current_stream(self.device).wait_stream(side)
            with torch.cuda.graph(self.graph, stream=side):
                self.out = model(self.x, self.t, return_dict=False, **kw)[0]
        self.in_channels = getattr(model, "in_channels", shape[1])
        self.out_channels = getattr(model, "out_channels", None)

    def __call__(self, sample, timestep, return_dict=True, uniform_timestep=False):
        self.x.copy_(_lib.require_cuda_f32(sample, "sample"))
        if torch.is_tensor(timestep):
            self.t.copy_(timestep.to(self.t.dtype).expand_as(self.t) if timestep.dim() == 0 else timestep.to(self.t.dtype))
        else:
            self.t.fill_(timestep)
        self.graph.replay()
        if not return_dict:
            return (self.out,)
        return SimpleNamespace(sample=self.out)


def _call_model_iadb(model, uniform_timestep=False):
    """``uniform_timestep``: every sample of a step carries the same UNet timestep (true for the reference's loops);
    evaluators that can exploit it (FusedUNet2D) are told so."""
    if uniform_timestep and getattr(model, "supports_uniform_timestep", False):
        return lambda x, t: model(x, t, return_dict=False, uniform_timestep=True)[0]
    return lambda x, t: model(x, t, return_dict=False)[0]


class IadbSampler:
    """A sampling run of fixed shape and schedule, set up once and callable many times: the
    coefficient table, the device-side step state, the static x buffer and (optionally) the
    captured [UNet -> K2] graph all persist across calls.  ``sample_iadb(..., use_graph=True)``
    keeps one of these per (model, shape, schedule).

    graph: None (eager), 'step' (UNet+K2 in one graph, replayed T times), 'unet' (UNet graph + eager K2) or 'loop' (all T
    steps in ONE graph, one replay per run; no per-step callback).
    time_step_kernel: record a CUDA event pair around every K2 launch (graph must not be 'step');
    ``step_kernel_ms()`` then returns the per-launch durations of the last run."""

    def __init__(self, model, shape, nb_step, scheduler_gamma="sigmoid", scheduler_params=(1000.0, 0.0, 3.0),
                 out_channel=6, noise_type="gaussianBN", scheduler_alpha="linear", alpha_param=1000.0, x_c=None,
                 device="cuda", graph="step", time_step_kernel=False, table=None, first_t=None, schedule_device="cpu",
                 schedule_nb_steps=None):
        self.device = torch.device(device)
        self.shape = tuple(shape)
        B, C = self.shape[0], self.shape[1]
        self.nb_step = nb_step
        if table is None:
            _expected_out_channels(noise_type, out_channel, C)
            table, first_t = iadb_table(nb_step, scheduler_alpha, scheduler_gamma,
                                        tuple(float(p) for p in scheduler_params), alpha_param, batch=B,
                                        device=schedule_device, schedule_nb_steps=schedule_nb_steps)
        else:
            table = table.clone()
        two = noise_type in TWO_HEAD and out_channel == 2 * C
        if not two:
            table[..., 1] = 0.0
        self.stepper = IadbStepper(table, first_t, B, self.device, expect_channels=2 * C if two else C)
        self.x = torch.zeros(self.shape, dtype=torch.float32, device=self.device)
        # the UNet timestep of a step is alpha_start of ONE t for the whole batch (iadb_bn.py:306-311); with the
        # non-linear alpha schedules torch's CPU vector / scalar paths may differ in the last ulp across the batch, so
        # the hint is only given when the table says the values really are identical
        t_first = torch.as_tensor(first_t, dtype=torch.float32).reshape(-1)
        uniform = bool((table[:, :, 2] == table[:, :1, 2]).all()) and bool((t_first == t_first[0]).all())
        call = _call_model_iadb(model, uniform)
        self.x_c = None
        if x_c is not None:
            # a STATIC conditioning buffer: every run copies its x_c in, so a captured graph never closes over a
            # caller tensor (new batch = new x_c = same graph) and never replays stale conditioning
            self.x_c = torch.empty(tuple(x_c.shape), dtype=torch.float32, device=self.device)
            self.x_c.copy_(x_c)
            inner, buf = call, self.x_c
            call = lambda xx, tt: inner(torch.cat([xx, buf], 1), tt)          # iadb_bn.py:406
        self.call = call
        try:
            self._model_ref = weakref.ref(model)
        except TypeError:
            self._model_ref = lambda: model
        if time_step_kernel and graph == "step":
            raise ValueError("time_step_kernel needs graph in (None, 'unet')")
        self.graphed = None
        self.looped = None
        if graph == "loop":
            if time_step_kernel:
                raise ValueError("time_step_kernel needs graph in (None, 'unet')")
            with torch.no_grad():
                self.looped = GraphedLoop(call, self.stepper, self.x, nb_step)
        elif graph is not None:
            with torch.no_grad():
                self.graphed = GraphedStep(call, self.stepper, self.x, fused=(graph == "step"))
        self.events = None
        if time_step_kernel:
            self.events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                           for _ in range(nb_step)]
        self.kernel_launches_per_run = nb_step            # K2 launches of one run (graph nodes count)

    @torch.no_grad()
    def run(self, x0, on_step=None, x_c=None):
        """x0 is copied into the static buffer (the reference never writes into x0,
        iadb_bn.py:326 makes new tensors); returns the static buffer (clone it to keep it)."""
        self.x.copy_(_lib.require_cuda_f32(x0, "x0"))
        if x_c is not None:
            if self.x_c is None:
                raise ValueError("this sampler was built without conditioning")
            self.x_c.copy_(x_c)
        self.stepper.reset()
        if self.looped is not None:
            if on_step is not None:
                raise ValueError("graph='loop' replays the whole run at once: no per-step callback (use graph='step')")
            self.looped.replay()
            return self.x
        for i, t in enumerate(reversed(range(self.nb_step))):
            ev = self.events[i] if self.events is not None else None
            if self.graphed is not None:
                self.graphed.replay(ev)
            else:
                d = self.call(self.x, self.stepper.t_vec)
                if ev is not None:
                    ev[0].record()
                self.stepper.step_(self.x, d)
                if ev is not None:
                    ev[1].record()
            if on_step is not None:
                on_step(t, self.x)
        return self.x

    def __call__(self, x0):
        return self.run(x0).clone()

    def step_kernel_ms(self):
        """Per-launch K2 durations (ms) of the last run; synchronises on the last event."""
        if self.events is None:
            raise ValueError("constructed without time_step_kernel")
        self.events[-1][1].synchronize()
        return [a.elapsed_time(b) for a, b in self.events]


_sampler_cache: "dict[tuple, IadbSampler]" = {}


def _run_iadb(model, x0, x_c, nb_step, scheduler_alpha, scheduler_gamma, scheduler_params, out_channel, noise_type,
              train_or_test, log_freq, use_graph, alpha_param=1000.0, schedule_device="cpu", schedule_nb_steps=None):
    x0 = _lib.require_cuda_f32(x0, "x0")
    params = tuple(float(p) for p in scheduler_params)
    key = (id(model), tuple(x0.shape), str(x0.device), nb_step, scheduler_alpha, scheduler_gamma, params, out_channel,
           noise_type, float(alpha_param), None if x_c is None else tuple(x_c.shape), str(schedule_device), schedule_nb_steps,
           "loop" if use_graph == "loop" else "step")
    sampler = _sampler_cache.get(key) if use_graph else None
    if sampler is not None and sampler._model_ref() is not model:      # id() reuse after garbage collection
        sampler = None
    if sampler is None:
        sampler = IadbSampler(model, x0.shape, nb_step, scheduler_gamma, params, out_channel, noise_type,
                              scheduler_alpha, alpha_param, x_c, x0.device,
                              graph=("loop" if use_graph == "loop" else "step") if use_graph else None,
                              schedule_device=schedule_device, schedule_nb_steps=schedule_nb_steps)
        if use_graph:
            if len(_sampler_cache) >= 4:
                _sampler_cache.pop(next(iter(_sampler_cache)))
            _sampler_cache[key] = sampler
    if nb_step == 1000:
        log_freq = 100
    if use_graph == "loop":
        if train_or_test == "test":
            raise ValueError("use_graph='loop' has no per-step snapshots: train_or_test must not be 'test'")
        return sampler.run(x0, x_c=x_c).clone(), [], []

    x_all, per_step = [], []
    tic = [time.time()]

    def on_step(t, x):
        per_step.append(time.time() - tic[0])
        if train_or_test == "test" and (t % log_freq == 0 or t == nb_step - 1):
            x_all.append(x.clone())
        tic[0] = time.time()
    x = sampler.run(x0, on_step, x_c=x_c).clone()
    return x, x_all, per_step


def sample_iadb(model, x0, nb_step, *args, use_graph=False, schedule_device="cpu", **kwargs):
    """Both reference signatures (see module docstring).  Test mode returns
    ``(x, x_all, mean_step_seconds)`` like iadb_bn.py:376-378, otherwise ``x``.
    ``use_graph``: False (eager), True (one captured [UNet -> K2] step replayed T times) or 'loop' (the whole T-step run
    as ONE graph replay; not in test mode, which snapshots x between steps).
    ``schedule_device``: where alpha / gamma are evaluated ('cpu' = the reference's CPU path, the default; pass
    ``x0.device`` to reproduce a reference that runs its schedule kernels on the GPU, see schedules.iadb_table)."""
    if len(args) + len(kwargs) == 1:                  # iadb_bn.py:287 -- (scheduler_params,) + module `opt`
        scheduler_params = args[0] if args else kwargs["scheduler_params"]
        o = opt
        cfg = dict(scheduler_alpha=o.scheduler_alpha, scheduler_gamma=o.scheduler_gamma,
                   out_channel=o.out_channel, noise_type=o.noise_type, train_or_test=o.train_or_test,
                   log_freq=25, alpha_param=getattr(o, "scheduler_param", 1000.0),
                   schedule_nb_steps=getattr(o, "nb_steps", nb_step))   # iadb_bn.py:107,165 divide by opt.nb_steps
    else:                                             # utils.py:180
        names = ("scheduler_gamma", "scheduler_params", "out_channel", "noise_type", "train_or_test", "scheduler_alpha")
        bound = dict(zip(names, args))
        bound.update(kwargs)
        bound.setdefault("scheduler_alpha", "linear")
        scheduler_params = bound.pop("scheduler_params")
        cfg = dict(bound, log_freq=1)
    x, x_all, per_step = _run_iadb(model, x0, None, nb_step, scheduler_params=scheduler_params, use_graph=use_graph,
                                   schedule_device=schedule_device, **cfg)
    if cfg["train_or_test"] == "test":
        return x, x_all, (float(np.mean(per_step[1:])) if len(per_step) > 1 else float("nan"))
    return x


def sample_iadb_conditional(model, x0, x_c, nb_step, scheduler_params, *, use_graph=False):
    """iadb_bn.py:385 -- UNet input is cat([x, x_c], 1); returns (x, x_all) in test mode."""
    o = opt
    x, x_all, _ = _run_iadb(model, x0, x_c, nb_step, o.scheduler_alpha, o.scheduler_gamma, scheduler_params,
                            o.out_channel, o.noise_type, o.train_or_test, 25, use_graph,
                            getattr(o, "scheduler_param", 1000.0), schedule_nb_steps=getattr(o, "nb_steps", nb_step))
    if o.train_or_test == "test":
        return x, x_all
    return x


# ------------------------------------------------------------------ latent IADB (config 5)
class IADBScheduler:
    """latent_iadb_bn_diffusers.py:75-138.  ``noise_type`` / ``out_channels`` are the
    ``args`` globals that file reads (:108-119)."""

    def __init__(self, num_train_timesteps: int = 1000, noise_type="gaussianBN", out_channels=8):
        self.num_train_timesteps = num_train_timesteps
        self.num_inference_steps = None
        self.noise_type = noise_type
        self.out_channels = out_channels

    def set_timesteps(self, num_inference_steps: int):
        self.num_inference_steps = num_inference_steps

    def step(self, model_output, timestep: int, x_alpha):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        C = x_alpha.shape[1]
        if self.noise_type in TWO_HEAD:
            if self.out_channels not in (C, 2 * C):
                raise NotImplementedError
        elif self.noise_type != "gaussian":
            raise NotImplementedError
        N = self.num_inference_steps
        d = (timestep + 1) / N - timestep / N                      # python float, as :99-103
        coef = torch.full((x_alpha.shape[0],), d, dtype=torch.float32, device=x_alpha.device)
        two = self.noise_type in TWO_HEAD and self.out_channels == 2 * C
        if not two and model_output.shape[1] != C:
            raise ValueError("model_output channel count does not match the configured out_channels")
        return iadb_step(x_alpha, model_output, coef, coef if two else None)

    def add_noise(self, original_samples, noise, alpha):
        """Forward blend (:127-138) -- training-side helper, plain torch."""
        return (1 - alpha).view(-1, 1, 1, 1) * original_samples + alpha.view(-1, 1, 1, 1) * noise


@torch.no_grad()
def sample_latent_iadb(model, noise, num_steps, noise_type="gaussianBN", out_channels=8, use_graph=False):
    """The latent sampling loop latent_iadb_bn_diffusers.py:524-534 (VAE decode left to the
    caller).  UNet timestep is alpha=(t+1)/N broadcast over the batch (:525-528)."""
    x = _lib.require_cuda_f32(noise, "noise")
    C = x.shape[1]
    if noise_type in TWO_HEAD:
        if out_channels not in (C, 2 * C):
            raise NotImplementedError
    elif noise_type != "gaussian":
        raise NotImplementedError
    # graph mode: the captured [UNet -> K2] step is kept per (model, shape, schedule) like sample_iadb's,
    # so repeated calls (a sampling job runs hundreds of batches) replay instead of re-capturing
    key = ("latent", id(model), tuple(x.shape), str(x.device), num_steps, noise_type, out_channels)
    sampler = _sampler_cache.get(key) if use_graph else None
    if sampler is not None and sampler._model_ref() is not model:      # id() reuse after garbage collection
        sampler = None
    if sampler is None:
        table, first_t = latent_table(num_steps, batch=x.shape[0])
        sampler = IadbSampler(model, x.shape, num_steps, out_channel=out_channels, noise_type=noise_type, device=x.device,
                              graph="step" if use_graph else None, table=table, first_t=first_t)
        if use_graph:
            if len(_sampler_cache) >= 4:
                _sampler_cache.pop(next(iter(_sampler_cache)))
            _sampler_cache[key] = sampler
    return sampler(x)
