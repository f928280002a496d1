"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed).

The path has no cross-sample coupling: every image -- and every 64x64 tile
column of the contraction -- is independent, so ranks own contiguous slices of the batch
and the step loop contains NO collective.  Collectives exist only at the edges:
one broadcast of ``cov_mat_L`` (64 MiB) at init and an optional all-gather of finished
images.  This replaces ``torch.nn.DataParallel`` (iadb_bn.py:573,716,838), which
re-broadcasts all UNet parameters and scatters/gathers activations on EVERY forward.

Seed-parity rule: draw the GLOBAL white field with the global seed on every rank and keep
this rank's slice, so an N-GPU run reproduces the 1-GPU run bit for bit.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world_size: int):
    """Contiguous split; the first (global_batch % world_size) ranks get one extra sample."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(global_batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world_size)
    return t[lo:hi]


def global_white_draw(shape, seed: int, rank: int, world_size: int, device=None) -> torch.Tensor:
    """np.random.seed(seed); np.random.randn(*shape) (iadb_bn.py:75,761) -> this rank's slice."""
    rs = np.random.RandomState(seed)
    z = rs.randn(*shape).astype(np.float32)
    lo, hi = shard_bounds(shape[0], rank, world_size)
    out = torch.from_numpy(z[lo:hi].copy())
    return out.to(device) if device is not None else out


def broadcast_L(cov_mat_L: torch.Tensor, src: int = 0) -> torch.Tensor:
    """The one init-time collective: rank `src` holds L (loaded from the .npz), the others
    pass an empty (4096,4096) buffer.  NCCL over NVLink/NVSwitch for CUDA tensors."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(cov_mat_L, src=src)
    return cov_mat_L


def broadcast_module(module: torch.nn.Module, src: int = 0) -> None:
    """Replicate (random-init or loaded) UNet weights once at init."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for p in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(p.data, src=src)


def gather_images(local: torch.Tensor, global_batch: int) -> torch.Tensor:
    """Terminal all-gather of finished images (ragged shards are padded to the largest)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(global_batch, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
