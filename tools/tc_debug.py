"""Diagnostic: tcgen05 contraction vs the SIMT fp32 kernel and an fp64 reference, per shape."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bndm_b200 as bb
from bndm_b200.synth import hashed_tril

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
L_np = hashed_tril(seed=0)
L = torch.from_numpy(L_np).to(dev)
L64 = L.double()
for B, C in [(4, 3), (1, 1), (5, 3), (64, 3), (100, 3), (16, 4)]:
    x = torch.randn(B, C, 64, 64, device=dev)
    ref = (x.double().reshape(B * C, 4096) @ L64.T).reshape(B, C, 64, 64)
    res = {}
    for gemm in ("simt", "tc"):
        try:
            bn = bb.get_noise_v2(dev, x, L, None, None, "GBN", "train", True, gemm=gemm)[1]
            torch.cuda.synchronize()
            err = (bn.double() - ref).abs()
            res[gemm] = bn
            print(f"B={B} C={C} {gemm}: max|err|={err.max().item():.3e} rms={err.pow(2).mean().sqrt().item():.3e} "
                  f"ref_rms={ref.pow(2).mean().sqrt().item():.3f}")
            if err.max().item() > 1e-3:
                e = err.reshape(B * C, 4096)
                bad_cols = (e.max(1).values > 1e-3).nonzero().flatten()[:16].tolist()
                bad_rows = (e.max(0).values > 1e-3).nonzero().flatten()
                print("   bad cols:", bad_cols, " bad rows: n=", bad_rows.numel(), bad_rows[:24].tolist())
                print("   got[0,:8]", bn.reshape(B * C, 4096)[0, :8].tolist())
                print("   ref[0,:8]", ref.reshape(B * C, 4096)[0, :8].tolist())
        except Exception as ex:
            print(f"B={B} C={C} {gemm}: EXC {type(ex).__name__}: {ex}")
            sys.exit(1)
# timing
h = bb.prepare_L(L)
h.profile(True)
for B in (4, 64):
    x = torch.randn(B, 3, 64, 64, device=dev); g = torch.rand(B, device=dev)
    for gemm in ("tc", "simt"):
        ts = []
        for it in range(12):
            bb.get_noise_v2(dev, x, h, g, None, "gaussianBN", "train", True, gemm=gemm)
            ts.append(h.last_ms())
        ts = np.array(ts[2:])
        print(f"timing B={B} {gemm}: pack/gemm/epi ms = {np.median(ts, 0)}")
