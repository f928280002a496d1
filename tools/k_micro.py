"""Micro driver for profiling: a handful of get_noise_v2 calls (K1a/K1b/K1c) and scheduled IADB
steps (K2) with an L2 flush before each, so `ncu` sees a short, representative launch list.
    python tools/k_micro.py [--B 4 64] [--iters 3] [--k2]
Prints CUDA-event timings too (whole call; not valid under ncu)."""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bndm_b200 as bb
from bndm_b200.sampler import IadbStepper
from bndm_b200.schedules import iadb_table
from bndm_b200.synth import hashed_tril

p = argparse.ArgumentParser()
p.add_argument("--B", type=int, nargs="+", default=[4, 64])
p.add_argument("--C", type=int, default=3)
p.add_argument("--res", type=int, default=64)
p.add_argument("--iters", type=int, default=3)
p.add_argument("--k2", action="store_true")
p.add_argument("--no-flush", action="store_true")
p.add_argument("--flush", choices=["read", "write"], default="read",
               help="read: a 512 MB reduction leaves L2 full of CLEAN lines (like ncu's cache control); write: a memset leaves it full of DIRTY lines that must be written back while the kernel streams")
args = p.parse_args()

dev = torch.device("cuda:0")
L = torch.from_numpy(hashed_tril(seed=0)).to(dev)
h = bb.prepare_L(L, max_columns=max(args.B) * args.C * (4 if args.res == 128 else 1))
flush_buf = torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device=dev)
flush_sink = torch.zeros((), dtype=torch.float32, device=dev)


class _Flush:
    def zero_(self):
        if args.flush == "write":
            flush_buf.zero_()
        else:
            torch.sum(flush_buf, dim=0, out=flush_sink)


flush = _Flush()


def timed(fn, n):
    ts = []
    for _ in range(n):
        if not args.no_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return ts


def graph_timed(fn, reps=10):
    """us per call from one CUDA graph of `reps` x [L2 flush, call] minus a graph of `reps` flushes:
    no per-call event overhead or quantisation, cold L2 for every call."""
    def capture(body):
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                flush.zero_()
                body()
        return g
    out = []
    for body in (fn, lambda: None):
        g = capture(body)
        g.replay(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / reps)
        out.append(statistics.median(ts))
    return out[0] - out[1]


for B in args.B:
    x = torch.randn(B, args.C, args.res, args.res, device=dev)
    g = torch.rand(B, device=dev)
    h.reserve(B * args.C * (4 if args.res == 128 else 1))
    if not os.environ.get("NO_GRAPH_TIMING"):
        for want in (("noise",), ("noise", "bn", "wn")):
            us = graph_timed(lambda: bb.get_noise_v2(dev, x, h, g, None, "gaussianBN", "train", True, want=want))
            alg = 4 * 4096 * 4097 // 2 + 4 * 4096 * B * args.C * (1 + len(want))
            print(f"get_noise B={B} C={args.C} res={args.res} outputs={len(want)}: {us:.2f} us per call (graph, L2 cold) "
                  f"= {alg / us / 1e3:.0f} GB/s algorithmic", flush=True)
    for want in (("noise",), ("noise", "bn", "wn")):
        ts = timed(lambda: bb.get_noise_v2(dev, x, h, g, None, "gaussianBN", "train", True, want=want), args.iters + 2)
        print(f"get_noise B={B} C={args.C} res={args.res} outputs={len(want)}: us per call = "
              f"{[round(t, 1) for t in ts]} median {statistics.median(ts[2:] or ts):.1f}", flush=True)

if args.k2:
    for B, C, Cd, HW in [(64, 3, 6, 4096), (32, 3, 6, 16384), (16, 4, 8, 4096)]:
        table, first_t = iadb_table(8, batch=B)
        st = IadbStepper(table, first_t, B, dev)
        x = torch.randn(B, C, int(HW ** 0.5), int(HW ** 0.5), device=dev)
        d = torch.randn(B, Cd, int(HW ** 0.5), int(HW ** 0.5), device=dev)
        ts = timed(lambda: st.step_(x, d), 6)
        nbytes = 4 * B * HW * (2 * C + Cd)
        med = statistics.median(ts[2:] or ts)
        print(f"K2 B={B} C={C} Cd={Cd} HW={HW}: us = {[round(t, 1) for t in ts]} median {med:.1f} "
              f"({nbytes / med / 1e3:.0f} GB/s with event overhead)", flush=True)
        d_cl = d.contiguous(memory_format=torch.channels_last)
        ts = timed(lambda: st.step_(x, d_cl), 6)
        med = statistics.median(ts[2:] or ts)
        print(f"K2 (d channels-last) B={B} C={C} Cd={Cd} HW={HW}: us = {[round(t, 1) for t in ts]} median {med:.1f} "
              f"({nbytes / med / 1e3:.0f} GB/s with event overhead)", flush=True)

if args.k2:
    # K3 (DDIM step, cfg 3: B=64, 3x64x64, eta = 0 and eta = 1 with a noise field) and K4 (uint8 NHWC)
    from bndm_b200.ddim import DDIMScheduler, ddim_step_raw
    from bndm_b200.io import to_uint8_nhwc
    sch = DDIMScheduler()
    sch.set_timesteps(100)
    B = 64
    x = torch.randn(B, 3, 64, 64, device=dev)
    eps = torch.randn_like(x)
    nz = torch.randn_like(x)
    t_vec = torch.zeros(B, device=dev)
    for eta, noise in ((0.0, None), (1.0, nz)):
        table = sch.coefficient_table(eta).to(dev)
        state = torch.zeros(2, dtype=torch.int32, device=dev)
        ts = timed(lambda: ddim_step_raw(x, x, eps, noise, table, state, t_vec), 6)
        nbytes = 4 * x.numel() * (3 + (noise is not None))
        med = statistics.median(ts[2:] or ts)
        print(f"K3 DDIM step B={B} eta={eta}: us = {[round(t, 1) for t in ts]} median {med:.1f} "
              f"({nbytes / med / 1e3:.0f} GB/s with event overhead)", flush=True)
    ts = timed(lambda: to_uint8_nhwc(x), 6)
    print(f"K4 to_uint8_nhwc B={B}: us = {[round(t, 1) for t in ts]}", flush=True)

# ---- K1b duration vs amount of work (544 / 2112 / 4096 stages per column block): overhead + slope
if os.environ.get("K1_SCALING"):
    from bndm_b200 import _lib
    h.profile(True)
    for B in args.B:
        for label, res, extra in (("res32 (544 stages)", 32, 0), ("res64 tri (2112)", 64, 0), ("res64 dense (4096)", 64, _lib.FORCE_DENSE)):
            x = torch.randn(B, args.C, res, res, device=dev)
            o = torch.empty_like(x)
            ts = []
            for it in range(8):
                if not args.no_flush:
                    flush.zero_()
                rc = _lib.load().bndm_get_noise_f32(h._h, _lib.ptr(x), None, _lib.ptr(o), None, None, B, args.C, res,
                                                    _lib.SRC_IMAGE | extra, _lib.current_stream(dev))
                _lib.check(rc, "get_noise")
                ts.append(h.last_ms()[1] * 1e3)
            print(f"K1b B={B} {label}: gemm us = {[round(t, 1) for t in ts[2:]]}", flush=True)
