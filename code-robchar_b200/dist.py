"""Multi-GPU sharding of the sweep: one process per GPU, controller blocks per rank.

Every (sigma, controller, draw) evaluation is independent and statistics are per (sigma,
controller); only ranking / Kendall tau needs all controllers' statistics.  So each rank runs the
fused sweep on its contiguous controller block and ONE all-gather assembles the [15][S][C]
statistics on every rank (NCCL over NVLink on GPUs; the same code runs under gloo for CPU tests of
the host logic with a caller-supplied compute function).  Philox counters use GLOBAL controller
indices (c_offset), so results are bit-identical for any world size.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as td


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `n` items for `rank`; the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_stats(local: torch.Tensor, C_total: int, group=None) -> torch.Tensor:
    """local: [K][S][C_local] on this rank -> [K][S][C_total] on every rank (controller axis gathered
    in rank order).  Uneven shards are padded to the largest block for the collective."""
    if not (td.is_available() and td.is_initialized()) or td.get_world_size(group) == 1:
        return local
    world = td.get_world_size(group)
    sizes = [shard_bounds(C_total, world, r)[1] - shard_bounds(C_total, world, r)[0] for r in range(world)]
    cmax = max(sizes)
    K, S = local.shape[0], local.shape[1]
    pad = torch.zeros((K, S, cmax), dtype=local.dtype, device=local.device)
    pad[:, :, :local.shape[2]] = local
    out = torch.empty((world * K, S, cmax), dtype=local.dtype, device=local.device)
    td.all_gather_into_tensor(out, pad.contiguous(), group=group)   # concatenated along dim 0 in rank order
    out = out.reshape(world, K, S, cmax)
    return torch.cat([out[r, :, :, :sizes[r]] for r in range(world)], dim=2).contiguous()


def sharded_rim_sweep(ctrl, sigmas, B: int, nspin: int, inspin: int, outspin: int, *, dkw_eps: float = 0.0,
                      seed: int = 0, model: int = 0, zz: bool = False, fused: bool = True, group=None,
                      compute_fn=None) -> torch.Tensor:
    """[15][S][C] statistics of the whole controller set, computed on this rank's block and
    all-gathered.  compute_fn(ctrl_block, c_offset) -> [15][S][C_local] overrides the device sweep
    (used by the gloo host-logic tests)."""
    ctrl = np.asarray(ctrl, dtype=np.float64) if not isinstance(ctrl, torch.Tensor) else ctrl
    C_total = ctrl.shape[0]
    world = td.get_world_size(group) if (td.is_available() and td.is_initialized()) else 1
    rank = td.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(C_total, world, rank)
    block = ctrl[lo:hi]
    if compute_fn is not None:
        local = compute_fn(block, lo)
    else:
        from . import engine
        if fused:
            local = engine.fidelity_stats(block, sigmas, B, nspin, inspin, outspin, dkw_eps=dkw_eps, seed=seed,
                                          model=model, zz=zz, c_offset=lo)
        else:
            f = engine.fidelity_mc(block, sigmas, B, nspin, inspin, outspin, seed=seed, model=model, zz=zz, c_offset=lo)
            local = engine.stats(f, dkw_eps)
    return all_gather_stats(local, C_total, group)
