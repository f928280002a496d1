"""Per-CTA timeline of the tcgen05 contraction kernel (debug time stamps, see bndm_debug_set_trace)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bndm_b200 as bb
from bndm_b200 import _lib
from bndm_b200.synth import hashed_tril

dev = torch.device("cuda:0")
L = torch.from_numpy(hashed_tril(seed=0)).to(dev)
h = bb.prepare_L(L, max_columns=192)
trace = torch.zeros(148 * 24, dtype=torch.int64, device=dev)
_lib.check(_lib.load().bndm_debug_set_trace(h._h, _lib.ptr(trace)), "trace")
flush_buf = torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device=dev)
flush_sink = torch.zeros((), dtype=torch.float32, device=dev)
WRITE_FLUSH = bool(os.environ.get("WRITE_FLUSH"))
for B in (4, 64):
    x = torch.randn(B, 3, 64, 64, device=dev); g = torch.rand(B, device=dev)
    for it in range(3):
        flush_buf.zero_() if WRITE_FLUSH else torch.sum(flush_buf, dim=0, out=flush_sink)
        trace.zero_()
        bb.get_noise_v2(dev, x, h, g, None, "gaussianBN", "train", True, want=("noise",))
        torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(148, 24).astype(np.float64)
    g0 = t[:, 0].min()
    print(f"B={B}: kernel span (globaltimer) = {(t[:, 7].max() - g0) / 1e3:.2f} us; CTA start spread = {(t[:, 0].max() - g0) / 1e3:.2f} us")
    clk = (t[:, 6] - t[:, 1])
    ghz = (clk / (t[:, 7] - t[:, 0])).mean()
    print(f"   SM clock ~ {ghz:.2f} GHz; per-CTA cycles: total mean {clk.mean():.0f} max {clk.max():.0f}")
    for name, a, b in (("init (barriers, TMEM alloc, sync)", 1, 2), ("init -> first operands landed", 2, 3),
                       ("first operands -> last MMA issued", 3, 4), ("last MMA issued -> epilogue done", 4, 5),
                       ("epilogue done -> exit", 5, 6)):
        d = t[:, b] - t[:, a]
        print(f"   {name:40s} mean {d.mean() / ghz / 1e3:6.2f} us   min {d.min() / ghz / 1e3:6.2f}   max {d.max() / ghz / 1e3:6.2f}")
    for name, k in (("producer: waiting for a free stage", 20), ("issuer: waiting for operands", 18), ("issuer: waiting for a drained TMEM buffer", 19),
                    ("issuer: operand wait + MMA issue + commit", 21),
                    ("converter: waiting for TMA", 17), ("converter: converting (incl. fence + arrive)", 16)):
        print(f"   {name:44s} mean {t[:, k].mean() / ghz / 1e3:6.2f} us total per CTA")
    if t[:, 13].max() > 0:
        m = t[:, 13] > 0
        print(f"   fused combine on {int(m.sum())} CTAs (last segment end of each): segments per tile {t[m, 14].min():.0f}..{t[m, 14].max():.0f}")
        for name, a, b in (("partial written -> fence done", 8, 9), ("fence -> ticket known (barriers + atomic)", 9, 10),
                           ("ticket -> acquire fence done", 10, 11), ("partial loads + adds (last batch)", 11, 12), ("emit (last batch)", 12, 13),
                           ("whole combine", 10, 13)):
            d = (t[m, b] - t[m, a]) / ghz / 1e3
            print(f"   {name:44s} mean {d.mean():6.2f} us   min {d.min():6.2f}   max {d.max():6.2f}")
