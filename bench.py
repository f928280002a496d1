#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configs.

metric   images/sec of the sampling run a config names (default --config 2 = configs[1]: 64x64 IADB,
         250 steps, cat_res64 UNet, random init, batch 64 per GPU, noise_type gaussianBN / out_channel 6,
         gamma sigmoid tau=1000 -- scripts/sampling/cat_res64_test.sh:5-9 of the reference), plus
         get_noise GB/s against the measured HBM peak.
step     ONE pass of the hot path over one batch: x0 = get_noise_v2(white, gamma(T)) (the reference's
         call at iadb_bn.py:770-775) followed by the sampling loop ([UNet forward -> update] x T,
         iadb_bn.py:304-344 / latent_iadb_bn_diffusers.py:524-534 / ddim_diffusers.py:672-683).
value    whole-job images/sec, white field already resident in HBM when the clock starts.
e2e      same metric through the public API (bb.get_noise_v2 + bb.sample_*) with HOST buffers: pinned
         host white field -> device, result images -> pinned host, every step.
roofline the kernel of libbndm_b200.so with the largest share of the step: the UNet's fused GroupNorm
         kernel K5 (every launch of one eager forward bracketed by CUDA events, right after the timed
         steps, same model / batch / stream).  `roofline_get_noise` = the contraction at cfg 1 (K1g) and
         at this config's shape (K1b), whole call, no events between its PDL-chained launches;
         `roofline_step` = the update kernel K2 / K3, timed live with CUDA events inside the timed region.
reference_gpu  the reference's own op sequence (oracle port of utils.sample_iadb / the latent / DDIM loop
         + get_noise_v2's torch ops, stock NCHW UNet, eager) on the same B200, reduced step count stated.

    python bench.py [--config 2|3|4|5] [--gpus N] [--steps K] [--warmup W]     # this repo's CUDA path
    python bench.py --impl reference [...]                                      # reference CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...                       # one rank per GPU, weak scaling

One JSON line on stdout (rank 0).  oracle/ is imported ONLY by the cpu_baseline / reference_gpu legs and by
--impl reference (the reference's own algorithm timed on the host cores / on the GPU as the bar to beat).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "images/s"
L_TRI_BYTES = 4 * 4096 * 4097 // 2         # 33 562 624: lower triangle of L incl. diagonal, fp32

# BASELINE.json configs[1..4] (configs[0] is the CPU plumbing case of the test-suite).  `batch` is per GPU
# (weak scaling): cfg 4 = 256 over 8 GPUs, cfg 5 = 128 over 8 GPUs.
CONFIGS = {
    2: dict(key="configs[1]", kind="iadb", res=64, C=3, out=6, T=250, batch=64, gamma=("sigmoid", (1000.0, 0.0, 3.0)),
            metric="images/sec at 64x64 IADB 250-step sampling",
            what="IADB sampling cat_res64 UNet (random-init, 113.7 M params), 250 steps, noise_type gaussianBN, out_channel 6, "
                 "alpha linear, gamma sigmoid tau=1000 (scripts/sampling/cat_res64_test.sh:5-9)"),
    3: dict(key="configs[2]", kind="ddim", res=64, C=3, out=3, T=100, batch=64, eta=1.0,
            metric="images/sec at 64x64 DDIM 100-step sampling with time-varying blue variance noise",
            what="DDIM sampling church_res64 UNet (random-init), 100 steps (ddim_diffusers.py:639-640, :672-683), eta = 1 with "
                 "variance_noise = get_noise_v2(gaussianBN, gamma(t)) every step (north-star composition, SURVEY 8d; parity "
                 "of the DDIM arithmetic is unpinned: diffusers is not vendored)"),
    4: dict(key="configs[3]", kind="iadb", res=128, C=3, out=6, T=250, batch=32, gamma=("sigmoid", (0.2, 0.0, 3.0)),
            metric="images/sec at 128x128 IADB 250-step sampling",
            what="IADB sampling cat_res128 UNet (random-init, 116.3 M params), 250 steps, batch 256 sharded over 8 GPUs "
                 "(32 per GPU), gamma sigmoid tau=0.2 (scripts/sampling/cat_res128_test.sh:4-6)"),
    5: dict(key="configs[4]", kind="latent", res=64, C=4, out=8, T=250, batch=16,
            metric="images/sec at 512x512 latent IADB (4x64x64 latents) 250-step sampling",
            what="latent IADB cat_res512 UNet on 4x64x64 latents (random-init), 250 steps, batch 128 sharded over 8 GPUs "
                 "(16 per GPU), VAE decode excluded (weights unavailable offline; "
                 "scripts/sampling/latent_iadb_cat_res512_test.sh:3-6)"),
}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--config", type=int, choices=sorted(CONFIGS), default=2, help="BASELINE.json config (2 = configs[1], the headline)")
    p.add_argument("--batch", type=int, default=None, help="images per GPU (default: the config's)")
    p.add_argument("--nb-steps", type=int, default=None, help="denoising steps per sampling run (default: the config's)")
    p.add_argument("--unet-dtype", choices=["fp32", "bf16"], default="fp32",
                   help="fp32 = the reference's numerics (cuDNN TF32 conv allowed, torch default)")
    p.add_argument("--unet", choices=["fused", "plain"], default="fused",
                   help="fused = channels-last evaluation with the K5/K6 kernels (bndm_b200.fused_unet); plain = the "
                        "stock PyTorch module (NCHW)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-reference-gpu", action="store_true")
    p.add_argument("--no-extras", action="store_true", help="skip the get_noise micro section")
    p.add_argument("--cpu-batch", type=int, default=8)
    p.add_argument("--cpu-steps", type=int, default=4)
    p.add_argument("--ref-gpu-steps", type=int, default=25, help="denoising steps the reference-on-GPU leg runs (extrapolated)")
    a = p.parse_args()
    cfg = CONFIGS[a.config]
    a.batch = a.batch or cfg["batch"]
    a.nb_steps = a.nb_steps or cfg["T"]
    return a


def ncu_traffic(key):
    """DRAM bytes (read + write) per launch of a kernel from the committed ncu --set full capture
    (profiles/traffic.json, written from the per-round ncu summaries), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="bndm_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [s.strip() for s in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                    power.append(float(parts[3]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------ workload pieces
def build_model(cfg):
    from bndm_b200.unet import get_latent_model, get_model
    if cfg["kind"] == "latent":
        return get_latent_model(512, cfg["out"])
    return get_model(cfg["C"], cfg["out"], cfg["res"])


def gamma_of(bb, torch, cfg, t, batch, T):
    """gamma(t) per sample (white fraction) with the config's schedule; the latent script uses linear gamma."""
    tt = torch.full((batch,), float(t))
    if cfg["kind"] == "iadb":
        return bb.get_scheduler_gamma(tt, cfg["gamma"][0], cfg["gamma"][1], T)
    return tt / T


def reference_loop(cfg, torch, osam, onoise, model, white, L, gamma_T, n_steps_run, T, dev):
    """The reference's op sequence for one batch, eager, on `dev`: returns (seconds for get_noise, seconds for
    n_steps_run denoising steps).  oracle port = one torch op per reference op (oracle/sampler.py, oracle/noise.py)."""
    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
    with torch.no_grad():
        sync()
        t0 = time.perf_counter()
        if cfg["kind"] == "ddim":
            x0 = white                                       # ddim_diffusers.py:667-669: x0 is the saved gaussian field
        else:
            x0 = onoise.get_noise_torch(dev, white, L, gamma_T, None, "gaussianBN", "test", True)[0]
        sync()
        t1 = time.perf_counter()
        if cfg["kind"] == "iadb":
            osam.sample_iadb_utils(model, x0, n_steps_run, cfg["gamma"][0], cfg["gamma"][1], cfg["out"], "gaussianBN", "train")
        elif cfg["kind"] == "latent":
            osam.latent_loop(model, x0, n_steps_run, "gaussianBN", cfg["out"])
        else:
            B = white.shape[0]

            def noise_fn(i, t, x):
                g = torch.full((B,), min(1.0, float(t) / 1000.0), device=dev)
                return onoise.get_noise_torch(dev, x, L, g, None, "gaussianBN", "test", False)[0]
            osam.ddim_loop(model, x0, n_steps_run, eta=cfg["eta"], noise_fn=noise_fn)
        sync()
        t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def cpu_reference_sample(cfg, batch, n_steps_run, T, threads, L_np, seed=0):
    """Times the reference's algorithm (oracle port) on the host cores for `batch` images and `n_steps_run` of the T
    denoising steps; returns (images/s extrapolated to the full run, detail)."""
    import numpy as np
    import torch
    from oracle import noise as onoise
    from oracle import sampler as osam
    torch.set_num_threads(threads)
    cache = cpu_reference_sample.__dict__
    if cache.get("key") != cfg["key"]:
        torch.manual_seed(0)
        cache["model"] = build_model(cfg).eval()
        cache["L"] = torch.from_numpy(L_np)
        cache["key"] = cfg["key"]
    model, L = cache["model"], cache["L"]
    rs = np.random.RandomState(seed)
    white = torch.from_numpy(rs.randn(batch, cfg["C"], cfg["res"], cfg["res"]).astype(np.float32))
    t_noise, t_loop = reference_loop(cfg, torch, osam, onoise, model, white, L, torch.ones(batch), n_steps_run, T,
                                     torch.device("cpu"))
    full = t_noise + t_loop * (T / n_steps_run)
    return batch / full, {"t_get_noise_s": t_noise, "t_per_denoise_step_s": t_loop / n_steps_run, "t_full_run_s": full,
                          "t_measured_s": t_noise + t_loop}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: the reference is Python scripts, nothing to compile or
    install) on all host cores.  A step = get_noise_v2 + `cpu_steps` denoising steps on `cpu_batch` images; `value` is
    that sample extrapolated linearly to the config's step count.  `config` describes what was RUN."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from bndm_b200.synth import blue_noise_L
    cfg = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    L_np = blue_noise_L()
    vals, det, measured = [], None, []
    for i in range(args.warmup + args.steps):
        v, det = cpu_reference_sample(cfg, args.cpu_batch, args.cpu_steps, args.nb_steps, threads, L_np, seed=i)
        if i >= args.warmup:
            vals.append(v)
            measured.append(det["t_measured_s"])
    value = statistics.mean(vals)
    sample = (f"each step: get_noise_v2 + {args.cpu_steps} of {args.nb_steps} denoising steps on {args.cpu_batch} images "
              f"(oracle port of the reference on torch-CPU fp32, {threads} threads); images/s = batch / (t_get_noise + "
              f"t_steps x {args.nb_steps}/{args.cpu_steps}), i.e. extrapolated linearly by {args.nb_steps / args.cpu_steps:g}x")
    line = {"impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * statistics.mean(measured),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{cfg['key']}: {cfg['what']} -- REFERENCE ARM, bounded sample: {args.cpu_batch} images x "
                                   f"{args.cpu_steps} of {args.nb_steps} steps per bench step on the host CPU",
                       "res": cfg["res"], "batch_per_gpu": args.cpu_batch, "nb_steps": args.nb_steps,
                       "nb_steps_run": args.cpu_steps, "extrapolation_factor": args.nb_steps / args.cpu_steps,
                       "device": f"host CPU, {threads} threads, torch-CPU fp32",
                       "unet_eval": "stock PyTorch module (NCHW, fp32) on the CPU",
                       "ms_per_step_note": "measured wall time of one bounded sample (not extrapolated)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "detail": det},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, cfg, where, world):
    return {"workload": f"{cfg['key']}: {cfg['what']}; batch={args.batch} per GPU; x0 = get_noise_v2(white, gamma(T)) with a "
                        f"synthetic blue-noise Cholesky factor L" + ("" if cfg["kind"] != "ddim" else " (variance noise only)"),
            "res": cfg["res"], "batch_per_gpu": args.batch, "global_batch": args.batch * world, "nb_steps": args.nb_steps,
            "device": where,
            "unet_dtype": ("fp32 weights/activations, cuDNN conv TF32 allowed (torch default, as the reference)"
                           if args.unet_dtype == "fp32" else "bf16 weights/activations"),
            "unet_eval": ("channels-last, GroupNorm+SiLU(+time-embedding/bias adds) and bias+residual fused in "
                          "libbndm_b200.so (K5/K6), convolutions = cuDNN" if (args.unet == "fused" and args.unet_dtype == "fp32")
                          else "stock PyTorch module"),
            "white_field": "global np.random.RandomState(seed).randn(global_batch, C, H, W) sliced per rank "
                           "(dist.global_white_draw: the seed-parity rule of SURVEY 8e)",
            "l2": "no explicit flush in the sampling loop: one denoising step streams >1 GB of UNet activations and "
                  "455 MB of weights through the 126 MB L2; the get_noise micro section flushes L2 between iterations"}


# ------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import bndm_b200 as bb
    from bndm_b200 import _lib
    from bndm_b200.dist import broadcast_L, broadcast_module, gather_images, global_white_draw
    from bndm_b200.synth import blue_noise_L
    from bndm_b200.unet import count_forward_flops

    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (bndm_b200 has no CPU fallback; use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    pk = peaks()
    B, T, C, RES, OUT = args.batch, args.nb_steps, cfg["C"], cfg["res"], cfg["out"]
    kind = cfg["kind"]
    n_cols = B * C * (4 if RES == 128 else 1)

    # ---- init: L built on rank 0 and broadcast once over NCCL; UNet weights replicated once
    L = torch.empty(4096, 4096, dtype=torch.float32, device=dev)
    L_np = None
    if rank == 0:
        L_np = blue_noise_L()
        L.copy_(torch.from_numpy(L_np))
    broadcast_L(L)
    handle = bb.prepare_L(L, max_columns=max(n_cols, 192))
    torch.manual_seed(0)
    model_plain = build_model(cfg).to(dev).eval()
    broadcast_module(model_plain)
    flops_per_image = count_forward_flops(model_plain, RES, RES)
    model = model_plain
    if args.unet_dtype == "bf16":
        model = model_plain.to(torch.bfloat16)
    elif args.unet == "fused":
        from bndm_b200.fused_unet import fuse_unet
        model = fuse_unet(model_plain)

    gamma_T = gamma_of(bb, torch, cfg, T, B, T).to(dev)                  # == 1 (pure white at t = T)
    sampler = None
    if kind == "iadb":
        sampler = bb.IadbSampler(model, (B, C, RES, RES), T, cfg["gamma"][0], cfg["gamma"][1], OUT, "gaussianBN", device=dev,
                                 graph="unet", time_step_kernel=True)
    n_total = args.warmup + args.steps
    # the GLOBAL white field of bench step i is drawn with seed 1234 + i and sliced per rank: rank 0's shard is the
    # same images at any N, so its checksum below must not change with the number of GPUs
    whites_host = [global_white_draw((world * B, C, RES, RES), 1234 + i, rank, world).pin_memory() for i in range(n_total)]
    whites_dev = [w.to(dev) for w in whites_host]
    out_host = torch.empty(B, C, RES, RES, dtype=torch.float32).pin_memory()

    ddim_gammas = None
    if kind == "ddim":
        sched = bb.DDIMScheduler()
        sched.set_timesteps(T)
        ddim_gammas = [torch.full((B,), min(1.0, float(t) / 1000.0), device=dev) for t in sched.timesteps]
        from bndm_b200.sampler import GraphedModel
        gmodel = GraphedModel(model, (B, C, RES, RES), device=dev, uniform_timestep=True)

        def blue_variance(i, t, x):
            return bb.get_noise_v2(dev, x, handle, ddim_gammas[i], None, "gaussianBN", "test", False, want=("noise",))[0]

    def run_sampling(x_white, public_api):
        if kind == "iadb":
            x0 = bb.get_noise_v2(dev, x_white, handle, gamma_T, None, "gaussianBN", "test", True, want=("noise",))[0]
            if public_api:
                return bb.sample_iadb(model, x0, T, cfg["gamma"][0], cfg["gamma"][1], OUT, "gaussianBN", "train", use_graph=True)
            return sampler.run(x0)
        if kind == "latent":
            x0 = bb.get_noise_v2(dev, x_white, handle, gamma_T, None, "gaussianBN", "test", True, want=("noise",))[0]
            return bb.sample_latent_iadb(model, x0, T, "gaussianBN", OUT, use_graph=True)
        return bb.sample_ddim(gmodel, x_white, T, eta=cfg["eta"], noise_fn=blue_variance)

    def step_resident(i):
        return run_sampling(whites_dev[i], False)

    def step_e2e(i):
        w = whites_host[i].to(dev, non_blocking=True)
        x = run_sampling(w, True)
        out_host.copy_(x, non_blocking=True)
        return x

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, per_step=None):
        """W warm-ups, then exactly K steps between CUDA events, max over ranks (ms)."""
        for i in range(args.warmup):
            fn(i)
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for i in range(args.warmup, n_total):
            last = fn(i)
            if per_step is not None:
                per_step()
        e1.record()
        fence()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), last

    clocks = ClockSampler(local)
    clocks.start()
    # ---- device-resident value, with live per-launch timing of the update kernel (K2: CUDA events around every launch)
    k2_ms = []

    def collect():
        # reading the event durations waits for this step's events only: the next step is not enqueued yet either
        # way (the stream is serial), so this adds no device idle time beyond one host round trip per T-launch step
        if sampler is not None:
            k2_ms.extend(sampler.step_kernel_ms())
    torch.cuda.profiler.start()          # `ncu --profile-from-start off` lists exactly the launches of these steps
    ms_res, x_last = timed(step_resident, collect)
    torch.cuda.profiler.stop()
    clock_info = clocks.stop()
    x_last = x_last.clone()

    # ---- seed parity evidence: rank 0's final images of the last timed step (identical at any N) and the gathered batch
    def sha(t):
        return hashlib.sha256(t.detach().cpu().numpy().tobytes()).hexdigest()[:16]
    checksum = {"rank0_shard_sha256_16": sha(x_last) if rank == 0 else None,
                "rank0_shard_sum": float(x_last.double().sum().item()) if rank == 0 else None}
    gathered = gather_images(x_last, world * B)
    if rank == 0:
        checksum["global_batch_sha256_16"] = sha(gathered)
        checksum["global_images"] = int(gathered.shape[0])
        checksum["note"] = ("final images of the last timed step; rank 0's shard is the first batch_per_gpu images of the "
                            "global white field, so rank0_shard_* must be the same at N = 1, 2, 4, 8")
    del gathered

    # ---- K5 / K6 / K7 (the UNet's glue kernels, the largest share of device time among the kernels
    # of libbndm_b200.so): inside the timed steps they run from a CUDA graph, so they are timed live
    # right after, on the same model / batch / stream: one eager forward with every launch
    # bracketed by CUDA events
    roofline_glue = None
    if args.unet == "fused" and args.unet_dtype == "fp32":
        from bndm_b200 import fused_unet as fu
        t_probe = torch.full((B,), 0.5, device=dev)
        with torch.no_grad():
            model(whites_dev[0], t_probe, return_dict=False)
            torch.cuda.synchronize(dev)
            fu.TIMING = []
            model(whites_dev[0], t_probe, return_dict=False)
            torch.cuda.synchronize(dev)
        recs, fu.TIMING = fu.TIMING, None
        by = {}
        for name, nbytes, e0, e1, _ in recs:
            d = by.setdefault(name, [0, 0.0, 0, 0.0, 0])
            ms = e0.elapsed_time(e1)
            d[0] += nbytes; d[1] += ms; d[2] += 1
            if nbytes >= 64 * 1024 * 1024:                      # the launches that are not latency-bound
                d[3] += ms; d[4] += nbytes
        k5 = by.get("K5", [0, 1e-9, 0, 0.0, 0])

        # The kernel's own duration: the SAME launches (same tensors, same order) replayed from one CUDA graph, as they run
        # inside the timed steps -- an event pair around an eager launch adds ~3 us to kernels that take 4-60 us.
        def graph_ms(launches, reps=5):
            if not launches:
                return None
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for l in launches:
                    l()
            torch.cuda.current_stream(dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for l in launches:
                    l()
            g.replay()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            e1.synchronize()
            return e0.elapsed_time(e1) / reps
        k5_all = [r[4] for r in recs if r[0] == "K5"]
        k5_big = [r[4] for r in recs if r[0] == "K5" and r[1] >= 64 * 1024 * 1024]
        ms_all, ms_big = graph_ms(k5_all), graph_ms(k5_big)
        ach = k5[0] / (ms_all * 1e-3) / 1e9 if ms_all else 0.0
        roofline_glue = {"kernel": "groupnorm_nhwc_cluster_kernel (K5: fused GroupNorm + SiLU + adds, NHWC)", "bound": "hbm",
                         "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                         "traffic": ncu_traffic("groupnorm_nhwc_cluster_kernel B=64 C=128 64x64"),
                         "traffic_note": "ncu DRAM bytes of ONE launch at the largest shape (algorithmic 268 435 456 B)",
                         "peak_source": pk["source"] + " (burst copy bandwidth)",
                         "bytes_per_forward": k5[0], "ms_per_forward": ms_all, "launches_timed": k5[2],
                         "achieved_large_launches": (k5[4] / (ms_big * 1e-3) / 1e9) if ms_big else None,
                         "eager_event_pairs": {"ms_per_forward": k5[1], "gbs": k5[0] / (k5[1] * 1e-3) / 1e9,
                                               "gbs_large_launches": (k5[4] / (k5[3] * 1e-3) / 1e9) if k5[3] > 0 else None},
                         "note": "all K5 launches of one forward (same tensors, same order) replayed from one CUDA graph, CUDA "
                                 "events around 5 replays; algorithmic bytes = read x once + write y once; 'large' = the "
                                 "launches moving >= 64 MiB; eager_event_pairs = the same launches timed one by one in an "
                                 "eager forward (each pair adds ~3 us)",
                         "others": {k: {"launches": v[2], "ms_per_forward": v[1], "gbs": v[0] / (v[1] * 1e-3) / 1e9}
                                    for k, v in by.items() if k != "K5"}}
        del recs, k5_all, k5_big

    # ---- end to end through the public API with host buffers
    ms_e2e, _ = timed(step_e2e)

    images = world * B * args.steps
    value = images / (ms_res / 1000.0)
    e2e_value = images / (ms_e2e / 1000.0)
    white_bytes = B * C * RES * RES * 4
    unet_tflops = flops_per_image * B * T / ((ms_res / args.steps) * 1e-3) / 1e12
    kpf = getattr(model, "kernels_per_forward", 0) or 0
    noise_calls = T if kind == "ddim" else 1
    noise_launches = (1 if n_cols <= 16 else 3) + (1 if (RES != 64) else 0)

    line = {"metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.unet_dtype == "fp32" else "bf16", "data": "synthetic",
            "config": workload_config(args, cfg, "B200", world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": white_bytes,
                    "d2h_bytes_per_step": white_bytes, "ms_per_step": ms_e2e / args.steps,
                    "api": {"iadb": "bb.get_noise_v2(...) + bb.sample_iadb(..., use_graph=True)",
                            "latent": "bb.get_noise_v2(...) + bb.sample_latent_iadb(..., use_graph=True)",
                            "ddim": "bb.sample_ddim(graphed model, ..., eta=1, noise_fn=bb.get_noise_v2)"}[kind]
                           + "; pinned host in/out"},
            "gpu_launches": args.steps * (noise_calls * noise_launches + T * (1 + kpf)),
            "gpu_launches_note": f"per step: {noise_calls} x get_noise_v2 ({noise_launches} launches: "
                                 f"{'K1g' if n_cols <= 16 else 'K1a pack + K1b contraction + K1c combine'}) + {T} x (update kernel + "
                                 f"{kpf} K5-K11 launches inside the UNet forward); cuDNN/cuBLAS/ATen kernels are not counted",
            "clocks": clock_info,
            "checksum": checksum,
            # `roofline` = the kernel of libbndm_b200.so with the largest share of the step's device time: K5 when the
            # fused UNet runs (19 % of the step in the ncu launch list, profiles/)
            "roofline": roofline_glue,
            "unet": {"gflop_per_image_forward": flops_per_image / 1e9, "achieved_tflops": unet_tflops,
                     "frac_of_bf16_sustained_peak": unet_tflops / pk["bf16_tflops_sustained"],
                     "images_per_s_ceiling_at_bf16_sustained_peak":
                         pk["bf16_tflops_sustained"] * 1e12 / (flops_per_image * T) * world}}
    if k2_ms:
        k2_mean = statistics.mean(k2_ms)
        k2_bytes = 4 * B * RES * RES * (C + OUT + C)
        k2_gbs = k2_bytes / (k2_mean * 1e-3) / 1e9
        line["roofline_step"] = {"kernel": "iadb_step_kernel (K2)", "bound": "hbm", "achieved": k2_gbs, "peak": pk["hbm_gbs"],
                                 "unit": "GB/s", "frac": k2_gbs / pk["hbm_gbs"], "traffic": ncu_traffic("iadb_step_kernel B=64"),
                                 "bytes_per_launch": k2_bytes, "ms_per_launch": k2_mean, "launches_timed": len(k2_ms),
                                 "note": "CUDA-event pair around every launch inside the timed steps: the pair costs as much as "
                                         "the ~4 us kernel; d was written by the UNet's last conv just before, so x/d are "
                                         "L2-resident (not a DRAM figure)"}

    if rank == 0 and not args.no_extras:
        micro = get_noise_micro(torch, bb, dev, L, handle, pk, [(64, 4, 3), (RES, B, C)])
        line["get_noise"] = micro
        line["roofline_get_noise"] = roofline_get_noise(micro, pk, (RES, B, C))
        if line["roofline"] is None:
            line["roofline"] = line["roofline_get_noise"]["this_config"]
        if sampler is not None:
            add_k2_graph_timing(torch, line, sampler, args, cfg, dev, pk)
    if rank == 0 and world == 1 and not args.no_reference_gpu:
        line["reference_gpu"] = reference_on_gpu(torch, args, cfg, model_plain, whites_dev[0], L, dev, value)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, det = cpu_reference_sample(cfg, args.cpu_batch, args.cpu_steps, T, threads, L_np)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"get_noise_v2 + {args.cpu_steps} of {T} denoising steps on {args.cpu_batch} "
                                          f"images (oracle port, torch-CPU fp32), extrapolated linearly to {T} steps",
                                "detail": det}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)
    return 0


def reference_on_gpu(torch, args, cfg, model_plain, white, L, dev, our_value):
    """The bar SURVEY 8d / BASELINE.md 3 name: the reference's own op sequence on the same B200 -- eager loop with the
    per-step coefficient tensors built by ~20 tiny kernels and a host->device copy (iadb_bn.py:304-316), 3-5 update
    launches, stock NCHW UNet, get_noise_v2 as expand + bmm -- run for `ref_gpu_steps` of the T steps (stated) and
    extrapolated linearly."""
    from oracle import noise as onoise
    from oracle import sampler as osam
    B, T = white.shape[0], args.nb_steps
    n_run = max(2, min(args.ref_gpu_steps, T))
    try:
        gT = torch.ones(B, device=dev)
        reference_loop(cfg, torch, osam, onoise, model_plain, white, L, gT, 2, T, dev)            # warm-up (cuDNN plans)
        t_noise, t_loop = reference_loop(cfg, torch, osam, onoise, model_plain, white, L, gT, n_run, T, dev)
        full = t_noise + t_loop * (T / n_run)
        v = B / full
        return {"value": v, "unit": UNIT, "what": "oracle port of the reference's eager loop (one torch op per reference op) + "
                "stock NCHW UNet + get_noise_v2's torch op sequence, on this B200, same batch",
                "steps_run": n_run, "of": T, "extrapolation_factor": T / n_run, "t_get_noise_s": t_noise,
                "t_per_denoise_step_s": t_loop / n_run, "t_full_run_s": full, "speedup_value_over_reference_gpu": our_value / v}
    except Exception as e:          # measurement extra only
        return {"error": repr(e)}


def add_k2_graph_timing(torch, line, sampler, args, cfg, dev, pk):
    """K2 without the per-launch event overhead (a CUDA-event pair around a ~4 us kernel costs as much as the kernel):
    one CUDA graph of 50 scheduled steps on the sampler's own buffers, x / d L2-resident as in the sampling loop."""
    B, C, RES, OUT = args.batch, cfg["C"], cfg["res"], cfg["out"]
    k2_bytes = 4 * B * RES * RES * (C + OUT + C)
    try:
        st = sampler.stepper
        st.reset()
        d_probe = torch.randn(B, OUT, RES, RES, device=dev)
        if args.unet == "fused" and args.unet_dtype == "fp32":
            d_probe = d_probe.contiguous(memory_format=torch.channels_last)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            st.step_(sampler.x, d_probe)
        torch.cuda.current_stream(dev).wait_stream(side)
        st.reset()
        with torch.cuda.graph(g, stream=side):
            for _ in range(50):
                st.step_(sampler.x, d_probe)
        ts = []
        for _ in range(5):
            st.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1) / 50)
        k2_graph_ms = statistics.median(ts)
        line["roofline_step"]["ms_per_launch_in_graph"] = k2_graph_ms
        line["roofline_step"]["achieved_in_graph"] = k2_bytes / (k2_graph_ms * 1e-3) / 1e9
        line["roofline_step"]["frac_in_graph"] = k2_bytes / (k2_graph_ms * 1e-3) / 1e9 / pk["hbm_gbs"]
    except Exception as e:          # measurement extra only
        line["roofline_step"]["in_graph_error"] = repr(e)


def roofline_get_noise(micro, pk, this_shape):
    """ONE number per shape for the contraction: algorithmic bytes / whole-call time from the L2-cold graph-timed micro
    section (no CUDA events between the PDL-chained launches, which would serialise what PDL overlaps)."""
    def entry(shape, kernel):
        res, B, C = shape
        m = micro[f"B{B}_C{C}_res{res}"]["ours_3_outputs_l2_cold"]
        return {"kernel": kernel, "bound": "hbm", "achieved": m["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": m["frac_of_hbm_peak"], "us_per_call": m["us"], "algorithmic_bytes": m["algorithmic_bytes"],
                "peak_source": pk["source"] + " (burst copy bandwidth)",
                "timing": "whole get_noise_v2 call (all its launches), L2 flushed by a 512 MB memset before every call, one CUDA "
                          "graph of 10 x [flush, call] minus the flushes"}
    res, B, C = this_shape
    n_cols = B * C * (4 if res == 128 else 1)
    out = {"cfg1": entry((64, 4, 3), "gemv_kernel<12> (K1g: TMA-streamed fp32 FFMA2 contraction, complete rows per CTA, one launch)"),
           "this_config": entry(this_shape, "gemv_kernel (K1g)" if n_cols <= 16 else
                                "gemm_tc_kernel (K1b: triangular L.z contraction, tcgen05 3xTF32) + K1a pack + K1c combine")}
    out["cfg1"]["traffic"] = ncu_traffic("gemv_kernel<12> B=4")
    out["this_config"]["traffic"] = ncu_traffic("gemm_tc_kernel<96,0,1> B=64") if this_shape == (64, 64, 3) else None
    return out


def get_noise_micro(torch, bb, dev, L, handle, pk, shapes):
    """get_noise_v2 on its own at cfg 1 (B=4) and at the bench config's shape, whole call, L2 flushed before
    every call (a 512 MB memset, i.e. a buffer 4x the L2 is written), next to the reference's
    torch op sequence (get_noise_recent.py:105-116: clone, view/permute, matmul,
    permute/contiguous, lerp) on the same GPU.  Timing: one CUDA graph of 10 x [flush, call]
    minus one graph of 10 flushes, CUDA events around the replays -- no per-call event overhead
    (a call is ~20 us, an event pair costs 2-4 us and is quantised to ~1 us)."""
    from oracle import noise as onoise
    out = {}
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    reps = 10

    def graph_us(fn):
        def capture(body):
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                body()                                  # lazy initialisation outside capture
            torch.cuda.current_stream(dev).wait_stream(side)
            with torch.cuda.graph(g, stream=side):
                for _ in range(reps):
                    flush.zero_()
                    body()
            return g
        res = []
        for body in (fn, lambda: None):
            g = capture(body)
            g.replay()
            torch.cuda.synchronize(dev)
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3 / reps)
            res.append(statistics.median(ts))
        return res[0] - res[1]

    was_profiling = handle.profile_enabled
    handle.profile(False)                               # no events between the PDL-chained kernels
    for res_, B, C in dict.fromkeys(shapes):
        x = torch.randn(B, C, res_, res_, device=dev)
        gamma = torch.rand(B, device=dev)
        n_cols = B * C * (4 if res_ == 128 else 1)
        handle.reserve(n_cols)
        r = {}
        legs = [("ours_3_outputs", 3, lambda: bb.get_noise_v2(dev, x, handle, gamma, None, "gaussianBN", "train", True)),
                ("ours_1_output", 1, lambda: bb.get_noise_v2(dev, x, handle, gamma, None, "gaussianBN", "train", True, want=("noise",))),
                ("torch_eager_reference_ops", 3, lambda: onoise.get_noise_torch(dev, x, L, gamma, None, "gaussianBN", "train", True))]
        if n_cols <= 16:
            legs.insert(2, ("ours_tcgen05_3_outputs", 3, lambda: bb.get_noise_v2(dev, x, handle, gamma, None, "gaussianBN", "train", True, gemm="tc")))
        for name, n_out, fn in legs:
            try:
                us = graph_us(fn)
            except Exception as e:                      # e.g. the reference's 128^2 draw is made on the CPU: not capturable
                r[name + "_l2_cold"] = {"error": repr(e)[:200]}
                continue
            alg = L_TRI_BYTES * (0.25 if res_ == 32 else 1.0) + 4 * 4096 * n_cols * (1 + n_out)
            r[name + "_l2_cold"] = {"us": us, "algorithmic_bytes": alg, "gbs": alg / us / 1e3,
                                    "frac_of_hbm_peak": alg / us / 1e3 / pk["hbm_gbs"]}
        out[f"B{B}_C{C}_res{res_}"] = r
    handle.profile(was_profiling)
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
