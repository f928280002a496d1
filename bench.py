#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric   images/sec of 64x64 IADB sampling, 250 steps (configs[1]: cat_res64 UNet, random
         init, batch 64 per GPU, noise_type gaussianBN / out_channel 6, gamma sigmoid
         tau=1000 -- scripts/sampling/cat_res64_test.sh:5-9 of the reference), plus
         get_noise GB/s against the measured HBM peak.
step     ONE pass of the hot path over one batch: x0 = get_noise_v2(white, gamma(T)) (the
         reference's call at iadb_bn.py:770-775) followed by the 250-step sample_iadb loop
         ([UNet forward -> K2 update] x 250, iadb_bn.py:304-344).
value    whole-job images/sec, white field already resident in HBM when the clock starts.
e2e      same metric through the public API (bb.get_noise_v2 + bb.sample_iadb) with HOST
         buffers: pinned host white field -> device, result images -> pinned host, every step.
roofline the kernel of libbndm_b200.so with the largest share of the step: the UNet's fused
         GroupNorm kernel K5 (every launch of one eager forward bracketed by CUDA events, right
         after the timed steps, same model / batch / stream); `roofline_get_noise` = the L.z
         contraction K1b and `roofline_step` = the IADB update K2, both timed live with CUDA
         events on their launch stream inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path
    python bench.py --impl reference [...]                              # reference CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...               # one rank per GPU, weak scaling

One JSON line on stdout (rank 0).  oracle/ is imported ONLY by the cpu_baseline leg and by
--impl reference (the reference's own algorithm timed on the host cores).
"""
from __future__ import annotations

import argparse
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

METRIC = "images/sec at 64x64 IADB 250-step sampling"
UNIT = "images/s"
RES, CH, OUT_CH = 64, 3, 6
GAMMA_PARAMS = (1000.0, 0.0, 3.0)          # scripts/sampling/cat_res64_test.sh:7
L_TRI_BYTES = 4 * 4096 * 4097 // 2         # 33 562 624: lower triangle of L incl. diagonal, fp32


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--batch", type=int, default=64, help="images per GPU (configs[1]: 64)")
    p.add_argument("--nb-steps", type=int, default=250, help="denoising steps per sampling run (configs[1]: 250)")
    p.add_argument("--unet-dtype", choices=["fp32", "bf16"], default="fp32",
                   help="fp32 = the reference's numerics (cuDNN TF32 conv allowed, torch default)")
    p.add_argument("--unet", choices=["fused", "plain"], default="fused",
                   help="fused = channels-last evaluation with the K5/K6 kernels (bndm_b200.fused_unet); plain = the "
                        "stock PyTorch module (NCHW)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extras", action="store_true", help="skip the get_noise micro section")
    p.add_argument("--cpu-batch", type=int, default=8)
    p.add_argument("--cpu-steps", type=int, default=4)
    return p.parse_args()


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


# ------------------------------------------------------------- reference CPU path (oracle)
def cpu_reference_sample(batch, n_steps_run, nb_steps_full, threads, L_np, seed=0):
    """Times the reference's algorithm (oracle port: get_noise_v2 + utils.sample_iadb with the
    same UNet architecture) on the host cores for `batch` images and `n_steps_run` of the
    `nb_steps_full` denoising steps; returns (images/s extrapolated to the full run, detail)."""
    import numpy as np
    import torch
    from oracle import noise as onoise
    from oracle import sampler as osam
    from bndm_b200.unet import get_model
    torch.set_num_threads(threads)
    cache = cpu_reference_sample.__dict__
    if "model" not in cache:
        torch.manual_seed(0)
        cache["model"] = get_model(CH, OUT_CH, RES).eval()
        cache["L"] = torch.from_numpy(L_np)
    model, L = cache["model"], cache["L"]
    rs = np.random.RandomState(seed)
    white = torch.from_numpy(rs.randn(batch, CH, RES, RES).astype(np.float32))
    gamma_T = torch.ones(batch)
    with torch.no_grad():
        t0 = time.perf_counter()
        x0 = onoise.get_noise_torch(torch.device("cpu"), white, L, gamma_T, None, "gaussianBN", "test", True)[0]
        t1 = time.perf_counter()
        osam.sample_iadb_utils(model, x0, n_steps_run, "sigmoid", GAMMA_PARAMS, OUT_CH, "gaussianBN", "train")
        t2 = time.perf_counter()
    t_noise, t_loop = t1 - t0, t2 - t1
    full = t_noise + t_loop * (nb_steps_full / n_steps_run)
    return batch / full, {"t_get_noise_s": t_noise, "t_per_denoise_step_s": t_loop / n_steps_run, "t_full_run_s": full}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from bndm_b200.synth import blue_noise_L
    threads = os.cpu_count() or 1
    L_np = blue_noise_L()
    vals, det = [], None
    for i in range(args.warmup + args.steps):
        v, det = cpu_reference_sample(args.cpu_batch, args.cpu_steps, args.nb_steps, threads, L_np, seed=i)
        if i >= args.warmup:
            vals.append(v)
    value = statistics.mean(vals)
    sample = (f"each step: get_noise_v2 + {args.cpu_steps} of {args.nb_steps} denoising steps on {args.cpu_batch} images "
              f"(oracle port of the reference on torch-CPU fp32), time extrapolated linearly to {args.nb_steps} steps")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.cpu_batch / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, "host CPU"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "detail": det},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, where):
    return {"workload": f"configs[1]: IADB sampling cat_res64 UNet (random-init, 113.7 M params), {args.nb_steps} steps, "
                        f"batch={args.batch} per GPU, noise_type gaussianBN, out_channel 6, alpha linear, gamma sigmoid "
                        f"tau=1000; x0 = get_noise_v2(white, gamma(T)) with a synthetic blue-noise Cholesky factor L",
            "res": RES, "batch_per_gpu": args.batch, "nb_steps": args.nb_steps, "device": where,
            "unet_dtype": ("fp32 weights/activations, cuDNN conv TF32 allowed (torch default, as the reference)"
                           if args.unet_dtype == "fp32" else "bf16 weights/activations"),
            "unet_eval": ("channels-last, GroupNorm+SiLU(+time-embedding/bias adds) and bias+residual fused in "
                          "libbndm_b200.so (K5/K6), convolutions = cuDNN" if (args.unet == "fused" and args.unet_dtype == "fp32")
                          else "stock PyTorch module"),
            "l2": "no explicit flush in the sampling loop: one denoising step streams >1 GB of UNet activations and "
                  "455 MB of weights through the 126 MB L2; the get_noise micro section flushes L2 between iterations"}


# ------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import bndm_b200 as bb
    from bndm_b200 import _lib
    from bndm_b200.dist import broadcast_L, broadcast_module
    from bndm_b200.synth import blue_noise_L
    from bndm_b200.unet import count_forward_flops, get_model

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
    B, T = args.batch, args.nb_steps

    # ---- init: L built on rank 0 and broadcast once over NCCL; UNet weights replicated once
    L = torch.empty(4096, 4096, dtype=torch.float32, device=dev)
    L_np = None
    if rank == 0:
        L_np = blue_noise_L()
        L.copy_(torch.from_numpy(L_np))
    broadcast_L(L)
    handle = bb.prepare_L(L, max_columns=B * CH)
    torch.manual_seed(0)
    model = get_model(CH, OUT_CH, RES).to(dev).eval()
    broadcast_module(model)
    flops_per_image = count_forward_flops(model, RES, RES)
    if args.unet_dtype == "bf16":
        model = model.to(torch.bfloat16)
    elif args.unet == "fused":
        from bndm_b200.fused_unet import fuse_unet
        model = fuse_unet(model)

    gamma_T = bb.get_scheduler_gamma(torch.full((B,), float(T)), "sigmoid", GAMMA_PARAMS, T).to(dev)   # == 1
    sampler = bb.IadbSampler(model, (B, CH, RES, RES), T, "sigmoid", GAMMA_PARAMS, OUT_CH, "gaussianBN", device=dev,
                             graph="unet", time_step_kernel=True)
    n_total = args.warmup + args.steps
    rs = np.random.RandomState(1234 + rank)
    whites_host = [torch.from_numpy(rs.randn(B, CH, RES, RES).astype(np.float32)).pin_memory() for _ in range(n_total)]
    whites_dev = [w.to(dev) for w in whites_host]
    out_host = torch.empty(B, CH, RES, RES, dtype=torch.float32).pin_memory()
    handle.profile(True)

    def step_resident(i):
        x0 = bb.get_noise_v2(dev, whites_dev[i], handle, gamma_T, None, "gaussianBN", "test", True, want=("noise",))[0]
        return sampler.run(x0)

    def step_e2e(i):
        w = whites_host[i].to(dev, non_blocking=True)
        x0 = bb.get_noise_v2(dev, w, handle, gamma_T, None, "gaussianBN", "test", True, want=("noise",))[0]
        x = bb.sample_iadb(model, x0, T, "sigmoid", GAMMA_PARAMS, OUT_CH, "gaussianBN", "train", use_graph=True)
        out_host.copy_(x, non_blocking=True)
        return x

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn):
        """W warm-ups, then exactly K steps between CUDA events, max over ranks (ms)."""
        for i in range(args.warmup):
            fn(i)
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.warmup, n_total):
            fn(i)
        e1.record()
        fence()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    clocks = ClockSampler(local)
    clocks.start()
    # ---- device-resident value, with live per-launch timing of K1 (profile hook) and K2 (events)
    for i in range(args.warmup):
        step_resident(i)
    fence()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1_ms, k2_ms = [], []
    e0.record()
    for i in range(args.warmup, n_total):
        step_resident(i)
        # reading the event durations waits for this step's events only: the next step is not
        # enqueued yet either way (the stream is serial), so this adds no device idle time beyond
        # one host round trip per 250-launch step
        k1_ms.append(handle.last_ms())
        k2_ms.extend(sampler.step_kernel_ms())
    e1.record()
    fence()
    ms_res = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_res, op=dist.ReduceOp.MAX)
    ms_res = float(ms_res.item())
    clock_info = clocks.stop()

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
        for name, nbytes, e0, e1 in recs:
            d = by.setdefault(name, [0, 0.0, 0, 0.0, 0])
            ms = e0.elapsed_time(e1)
            d[0] += nbytes; d[1] += ms; d[2] += 1
            if nbytes >= 64 * 1024 * 1024:                      # the launches that are not latency-bound
                d[3] += ms; d[4] += nbytes
        k5 = by.get("K5", [0, 1e-9, 0, 0.0, 0])
        roofline_glue = {"kernel": "groupnorm_nhwc_cluster_kernel (K5: fused GroupNorm + SiLU + adds, NHWC)", "bound": "hbm",
                         "achieved": k5[0] / (k5[1] * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": k5[0] / (k5[1] * 1e-3) / 1e9 / pk["hbm_gbs"],
                         "traffic": ncu_traffic("groupnorm_nhwc_cluster_kernel B=64 C=128 64x64"),
                         "traffic_note": "ncu DRAM bytes of ONE launch at the largest shape (algorithmic 268 435 456 B)",
                         "peak_source": pk["source"] + " (burst copy bandwidth)",
                         "bytes_per_forward": k5[0], "ms_per_forward": k5[1], "launches_timed": k5[2],
                         "achieved_large_launches": (k5[4] / (k5[3] * 1e-3) / 1e9) if k5[3] > 0 else None,
                         "note": "all K5 launches of one eager forward (B=64), algorithmic bytes = read x once + write y "
                                 "once; each launch carries ~3 us of event overhead; 'large' = launches moving >= 64 MiB",
                         "others": {k: {"launches": v[2], "ms_per_forward": v[1], "gbs": v[0] / (v[1] * 1e-3) / 1e9}
                                    for k, v in by.items() if k != "K5"}}

    # ---- end to end through the public API with host buffers
    ms_e2e = timed(step_e2e)

    images = world * B * args.steps
    value = images / (ms_res / 1000.0)
    e2e_value = images / (ms_e2e / 1000.0)
    white_bytes = B * CH * RES * RES * 4

    # ---- rooflines (rank 0's launches)
    n_cols = B * CH
    gemm_ms = statistics.mean(m[1] for m in k1_ms)
    k1_bytes = L_TRI_BYTES + 4 * 4096 * n_cols * (1 + 1)          # L triangle + read z + write `noise` (1 output)
    k1_flops = 2 * n_cols * (4096 * 4097 // 2)
    k1_gbs = k1_bytes / (gemm_ms * 1e-3) / 1e9
    k2_mean = statistics.mean(k2_ms)
    k2_bytes = 4 * B * RES * RES * (CH + OUT_CH + CH)
    k2_gbs = k2_bytes / (k2_mean * 1e-3) / 1e9
    tf32_peak = pk["bf16_tflops"] / 2.0
    roofline = {"kernel": "gemm_tc_kernel (K1b: triangular L.z contraction, tcgen05 3xTF32)", "bound": "hbm",
                "achieved": k1_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": k1_gbs / pk["hbm_gbs"],
                "traffic": ncu_traffic("gemm_tc_kernel<96,0,1> B=64"), "peak_source": pk["source"] + " (burst copy bandwidth)",
                "bytes_per_launch": k1_bytes, "ms_per_launch": gemm_ms, "launches_timed": len(k1_ms),
                "pack_ms": statistics.mean(m[0] for m in k1_ms), "epilogue_ms": statistics.mean(m[2] for m in k1_ms),
                "tensor": {"achieved_tflops_fp32_equiv": k1_flops / (gemm_ms * 1e-3) / 1e12,
                           "achieved_tflops_tf32_issued": 3 * k1_flops / (gemm_ms * 1e-3) / 1e12,
                           "peak_tf32_tflops": tf32_peak, "peak_note": "measured bf16 cuBLAS burst / 2",
                           "frac_issued": 3 * k1_flops / (gemm_ms * 1e-3) / 1e12 / tf32_peak}}
    roofline_step = {"kernel": "iadb_step_kernel (K2)", "bound": "hbm", "achieved": k2_gbs, "peak": pk["hbm_gbs"],
                     "unit": "GB/s", "frac": k2_gbs / pk["hbm_gbs"], "traffic": ncu_traffic("iadb_step_kernel B=64"),
                     "bytes_per_launch": k2_bytes,
                     "ms_per_launch": k2_mean, "launches_timed": len(k2_ms),
                     "note": "d was written by the UNet's last conv just before: x/d are L2-resident, so this can "
                             "exceed the DRAM copy peak"}
    unet_tflops = flops_per_image * B * T / ((ms_res / args.steps) * 1e-3) / 1e12

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.unet_dtype == "fp32" else "bf16", "data": "synthetic",
            "config": workload_config(args, "B200"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": white_bytes,
                    "d2h_bytes_per_step": white_bytes, "ms_per_step": ms_e2e / args.steps,
                    "api": "bb.get_noise_v2(...) + bb.sample_iadb(..., use_graph=True); pinned host in/out"},
            "gpu_launches": args.steps * (3 + T * (1 + (getattr(model, "kernels_per_forward", 0) or 0))),
            "gpu_launches_note": f"per step: K1a pack + K1b contraction + K1c combine + {T} x (K2 + "
                                 f"{getattr(model, 'kernels_per_forward', 0) or 0} K5/K6 launches inside the UNet forward); "
                                 "cuDNN/cuBLAS/ATen kernels are not counted",
            "clocks": clock_info,
            # `roofline` = the kernel of libbndm_b200.so with the largest share of the step's device time: K5 when the
            # fused UNet runs (19 % of the step in the ncu launch list, profiles/), else the contraction K1b
            "roofline": roofline_glue if roofline_glue is not None else roofline,
            "roofline_get_noise": roofline, "roofline_step": roofline_step,
            "unet": {"gflop_per_image_forward": flops_per_image / 1e9, "achieved_tflops": unet_tflops,
                     "frac_of_bf16_sustained_peak": unet_tflops / pk["bf16_tflops_sustained"],
                     "images_per_s_ceiling_at_bf16_sustained_peak":
                         pk["bf16_tflops_sustained"] * 1e12 / (flops_per_image * T) * world}}

    if rank == 0 and not args.no_extras:
        line["get_noise"] = get_noise_micro(torch, bb, dev, L, handle, pk)
        # K2 without the per-launch event overhead (a CUDA-event pair around a ~4 us kernel costs as much as
        # the kernel): one CUDA graph of 50 scheduled steps on the sampler's own buffers, x / d L2-resident
        # as in the sampling loop
        try:
            st = sampler.stepper
            st.reset()
            d_probe = torch.randn(B, OUT_CH, RES, RES, device=dev)
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
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, det = cpu_reference_sample(args.cpu_batch, args.cpu_steps, T, threads, L_np)
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


def get_noise_micro(torch, bb, dev, L, handle, pk):
    """get_noise_v2 on its own at cfg 1 (B=4) and cfg 2 (B=64), whole call, L2 flushed before
    every call (a 512 MB memset, i.e. a buffer 4x the L2 is written), next to the reference's
    torch op sequence (get_noise_recent.py:105-116: clone, view/permute, matmul,
    permute/contiguous, lerp) on the same GPU.  Timing: one CUDA graph of 10 x [flush, call]
    minus one graph of 10 flushes, CUDA events around the replays -- no per-call event overhead
    (a call is ~20 us, an event pair costs 2-4 us and is quantised to ~1 us)."""
    out = {}
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    reps = 10

    def torch_eager(x, gamma):
        noise = x
        wn = noise.clone()
        B, C = x.shape[0], x.shape[1]
        n = noise.view(B, C, -1).permute(0, 2, 1)
        bn = torch.matmul(L, n).permute(0, 2, 1).contiguous().view(B, C, RES, RES)
        return bn * (1 - gamma.view(-1, 1, 1, 1)) + wn * gamma.view(-1, 1, 1, 1), bn, wn

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
    for B in (4, 64):
        x = torch.randn(B, CH, RES, RES, device=dev)
        gamma = torch.rand(B, device=dev)
        n_cols = B * CH
        handle.reserve(n_cols)
        res = {}
        for name, n_out, fn in (("ours_3_outputs", 3, lambda: bb.get_noise_v2(dev, x, handle, gamma, None, "gaussianBN", "train", True)),
                                ("ours_1_output", 1, lambda: bb.get_noise_v2(dev, x, handle, gamma, None, "gaussianBN", "train", True, want=("noise",))),
                                ("torch_eager_reference_ops", 3, lambda: torch_eager(x, gamma))):
            us = graph_us(fn)
            alg = L_TRI_BYTES + 4 * 4096 * n_cols * (1 + n_out)
            res[name + "_l2_cold"] = {"us": us, "algorithmic_bytes": alg, "gbs": alg / us / 1e3,
                                      "frac_of_hbm_peak": alg / us / 1e3 / pk["hbm_gbs"]}
        out[f"B{B}_C{CH}_res{RES}"] = res
    handle.profile(was_profiling)
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
