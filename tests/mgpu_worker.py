"""Worker of tests/test_multi_gpu.py: one process per rank (torchrun).  Seed-parity rule of SURVEY 8e on hardware:

    global_white_draw -> per-rank get_noise_v2(shard) + sample_iadb -> gather_images

must equal, bit for bit, the same global batch processed shard by shard on ONE GPU (a shard's result does not depend
on which GPU computed it or on what the other ranks do), and must agree with the unsharded single-call result to
rounding (another batch size is another cuDNN / contraction instance, i.e. another summation order).
With >= world GPUs: NCCL, one GPU per rank.  With fewer (a 1-GPU box): every rank drives cuda:0 and the collectives go
through gloo on host copies -- still the CUDA kernels, still separate processes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bndm_b200 as bb  # noqa: E402
from bndm_b200.dist import broadcast_L, broadcast_module, gather_images, global_white_draw, shard_bounds  # noqa: E402
from bndm_b200.fused_unet import fuse_unet  # noqa: E402
from bndm_b200.synth import hashed_tril  # noqa: E402
from bndm_b200.unet import get_latent_model  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    n_gpu = torch.cuda.device_count()
    nccl = n_gpu >= world
    dev = torch.device("cuda", rank if nccl else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if nccl else "gloo", **({"device_id": dev} if nccl else {}))

    def coll(t):                       # tensor the collectives may touch
        return t if nccl else t.cpu()

    # ---- init: L and the UNet weights exist on rank 0 only and are broadcast once
    L = torch.empty(4096, 4096, dtype=torch.float32, device=dev)
    if rank == 0:
        L.copy_(torch.from_numpy(hashed_tril(seed=0)))
    Lc = coll(L)
    broadcast_L(Lc)
    L = Lc.to(dev)
    torch.manual_seed(rank)            # different init per rank on purpose: the broadcast must make them equal
    model = get_latent_model(256, 8).eval().to(dev if nccl else "cpu")
    broadcast_module(model)
    model = model.to(dev)
    fused = fuse_unet(model)

    G, C, RES, T, seed = 6, 4, 32, 6, 7
    gamma_g = torch.linspace(0.2, 1.0, G)

    def pipeline(lo, hi, white_global):
        """what a rank does for samples [lo, hi) of the global batch"""
        x0 = bb.get_noise_v2(dev, white_global, L, gamma_g.to(dev), None, "gaussianBN", "test", True, shard=(lo, hi))[0]
        return bb.sample_latent_iadb(fused, x0, T, "gaussianBN", 8, use_graph=True)

    white_global = global_white_draw((G, C, RES, RES), seed, 0, 1, device=dev)        # every rank draws the global field
    lo, hi = shard_bounds(G, rank, world)
    mine = global_white_draw((G, C, RES, RES), seed, rank, world, device=dev)
    assert torch.equal(mine, white_global[lo:hi])
    out = pipeline(lo, hi, white_global)
    gathered = gather_images(coll(out), G)

    # ---- 128^2 inplace get_noise: the (b', k') = divmod(k*B + b, 4) mixing needs the GLOBAL batch
    x128 = global_white_draw((4, 3, 128, 128), seed + 1, 0, 1, device=dev)
    g128 = torch.linspace(0.1, 0.9, 4, device=dev)
    lo2, hi2 = shard_bounds(4, rank, world)
    n128 = bb.get_noise_v2(dev, x128, L, g128, None, "gaussianBN", "test", True, shard=(lo2, hi2))
    g_noise, g_wn = gather_images(coll(n128[0]), 4), gather_images(coll(n128[2]), 4)

    ok = True
    if rank == 0:
        # the same global batch, shard by shard, on this one GPU
        seq = torch.cat([pipeline(*shard_bounds(G, r, world), white_global) for r in range(world)], 0)
        ok &= bool(torch.equal(gathered.to(dev), seq))
        whole = pipeline(0, G, white_global)
        err = (gathered.to(dev) - whole).abs().max().item()
        ok &= err < 1e-3 * max(1.0, whole.abs().max().item())
        full128 = bb.get_noise_v2(dev, x128, L, g128, None, "gaussianBN", "test", True)
        ok &= bool(torch.equal(g_wn.to(dev), full128[2]))
        e128 = (g_noise.to(dev) - full128[0]).abs().max().item()
        ok &= e128 < 2e-5
        seq128 = torch.cat([bb.get_noise_v2(dev, x128, L, g128, None, "gaussianBN", "test", True, shard=shard_bounds(4, r, world))[0]
                            for r in range(world)], 0)
        ok &= bool(torch.equal(g_noise.to(dev), seq128))
        print(f"MGPU backend={'nccl' if nccl else 'gloo'} gpus={n_gpu} world={world} sharded_vs_sequential_equal="
              f"{torch.equal(gathered.to(dev), seq)} sharded_vs_whole_maxerr={err:.3e} noise128_maxerr={e128:.3e} ok={ok}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev if nccl else "cpu")
    dist.broadcast(flag, src=0)
    ok = bool(flag.item())
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
