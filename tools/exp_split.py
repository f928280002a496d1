"""Experiment: how does tcgen05 kind::tf32 read a full-mantissa fp32 operand, and where does the
3xTF32 error come from?  Runs the tcgen05 contraction with BNDM_L_SPLIT_MODE = 0..3 (see
noise_pack.cu) on a full-mantissa random lower-triangular L and prints the error against fp64.
Usage: python tools/exp_split.py            (spawns one subprocess per mode)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    import torch
    import bndm_b200 as bb
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1)
    L = torch.tril(torch.randn(4096, 4096, generator=g) / 45.0).to(dev)      # rows have ~unit norm near the end
    L64 = L.double()
    for B, C in [(4, 3), (64, 3), (100, 3)]:
        x = torch.randn(B, C, 64, 64, generator=g).to(dev)
        ref = (x.double().reshape(B * C, 4096) @ L64.T).reshape(B, C, 64, 64)
        for gemm in ("tc", "simt"):
            bn = bb.get_noise_v2(dev, x, L, None, None, "GBN", "train", True, gemm=gemm)[1]
            e = bn.double() - ref
            bias = (torch.sign(ref) * e).mean().item()
            print(f"mode={os.environ.get('BNDM_L_SPLIT_MODE', '0')} B={B} {gemm}: max={e.abs().max().item():.3e} "
                  f"rms={e.pow(2).mean().sqrt().item():.3e} bias_toward_sign={bias:+.3e} ref_rms={ref.pow(2).mean().sqrt().item():.3f}",
                  flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for mode in ("0", "1", "2", "3"):
            env = dict(os.environ, BNDM_L_SPLIT_MODE=mode)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, timeout=300)
