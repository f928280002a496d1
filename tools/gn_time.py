#!/usr/bin/env python
"""Whole-call timing of get_noise_v2 per kernel choice, L2 flushed (512 MB memset) before every call:
one CUDA graph of `reps` x [flush, call] minus one graph of `reps` flushes (bench.py's method).

    python tools/gn_time.py                       # default shape list, kernels auto / tc / gemv
    BNDM_GV_VARIANT=1 python tools/gn_time.py     # K1g: 4 quads per warp, scalar FFMA
Prints one line per (shape, kernel): us per call, algorithmic GB/s, fraction of the measured HBM peak, and for K1g
the per-CTA timeline from the trace hook (first stage landed / stream done / kernel end, mean and max over CTAs)."""
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bndm_b200 as bb  # noqa: E402
from bndm_b200 import _lib  # noqa: E402
from bndm_b200.synth import blue_noise_L  # noqa: E402

L_TRI = 4 * 4096 * 4097 // 2


def graph_us(dev, fn, flush, reps=10, kind="write"):
    def capture(body):
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream(dev).wait_stream(side)
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                if kind == "write":
                    flush.zero_()
                elif kind == "read":
                    flush.sum()
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


def main():
    dev = torch.device("cuda:0")
    peak = 6531.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    L = torch.from_numpy(blue_noise_L()).to(dev)
    h = bb.prepare_L(L, max_columns=512)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    shapes = [(64, 4, 3), (64, 4, 4), (64, 1, 3), (32, 4, 4), (64, 16, 4), (64, 64, 3), (128, 32, 3)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    kinds = os.environ.get("GN_FLUSH", "write,read").split(",")
    for res, B, C in shapes:
        x = torch.randn(B, C, res, res, device=dev)
        gamma = torch.rand(B, device=dev)
        n_cols = B * C * (4 if res == 128 else 1)
        h.reserve(n_cols)
        ref = None
        for gemm in ("auto", "tc", "gemv"):
            if gemm == "gemv" and n_cols > 16:
                continue
            for n_out, want in ((3, ("noise", "bn", "wn")), (1, ("noise",))):
                fn = lambda: bb.get_noise_v2(dev, x, h, gamma, None, "gaussianBN", "train", True, gemm=gemm, want=want)
                out = fn()[0]
                if ref is None:
                    ref = (x.double().reshape(-1, 4096) @ L.double().T) if res == 64 else None
                err = None
                if res == 64:
                    g4 = gamma.double().view(-1, 1, 1, 1)
                    want_out = ref.reshape(B, C, 64, 64) * (1 - g4) + x.double() * g4
                    err = (out.double() - want_out).abs().max().item()
                alg = L_TRI * (0.25 if res == 32 else 1.0) + 4 * 4096 * n_cols * (1 + n_out)
                for kind in kinds:
                    us = graph_us(dev, fn, flush, kind=kind)
                    print(f"res{res} B{B} C{C} N{n_cols:4d} {gemm:5s} outputs={n_out} flush={kind:5s}: {us:7.2f} us  "
                          f"{alg / us / 1e3:7.1f} GB/s  frac {alg / us / 1e3 / peak:.3f}  max|err| {err}")
        if n_cols <= 16:
            tr = torch.zeros(148 * 128 + 148 * 32, dtype=torch.int64, device=dev)
            _lib.check(_lib.load().bndm_debug_set_trace(h._h, _lib.C.c_void_p(tr.data_ptr())), "trace")
            for _ in range(3):
                flush.zero_()
                tr.zero_()
                bb.get_noise_v2(dev, x, h, gamma, None, "gaussianBN", "train", True, gemm="gemv")
                torch.cuda.synchronize()
            _lib.check(_lib.load().bndm_debug_set_trace(h._h, None), "trace")
            dd = tr.cpu().numpy()[148 * 128:].reshape(148, 8, 4).astype(np.float64)
            t = tr.cpu().numpy()[:148 * 128].reshape(148, 128).astype(np.float64)
            t0 = t[:, 0].min()
            for name, col in (("cta start", 0), ("first stage landed", 1), ("stream consumed", 2), ("outputs stored", 3)):
                v = (t[:, col] - t0) / 1e3
                print(f"    K1g {name:20s}: mean {v.mean():6.2f}  min {v.min():6.2f}  max {v.max():6.2f} us")
            if os.environ.get("GN_STAGES"):
                print("    stage: requested / landed / released (us after kernel start, mean over CTAs that have the stage)")
                for c in range(40):
                    m = t[:, 8 + c] > 0
                    if not m.any():
                        break
                    rq, ld, rl = ((t[m, col + c] - t0) / 1e3 for col in (8, 48, 88))
                    print(f"    {c:3d} ({int(m.sum()):3d} CTAs): {rq.mean():6.2f} {ld.mean():6.2f} {rl.mean():6.2f}   land-req {np.mean(ld - rq):5.2f}  rel-land {np.mean(rl - ld):5.2f}")


if __name__ == "__main__":
    main()
