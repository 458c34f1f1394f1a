"""Batch sharding of the ODE head across the GPUs of one node (one process per GPU, torch.distributed).

Samples are independent (reference loop future_prediction_ode.py:36-51), so the data path needs NO collective:
rank r integrates the contiguous sample range shard_bounds(B, r, P).  To stay stream-identical to the reference's single
sample-major noise stream, a rank first skips the noise draws that belong to the samples before its range.  An optional
all_gather assembles the full [B, T, C, H, W] output when one consumer needs it.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .schedule import merge_observations, plan_sample


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first n % world ranks get one extra sample."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def noise_draws_before(module, camera_timestamp, lidar_timestamp, target_timestamp, lidar_present: bool, lo: int, hi=None) -> int:
    """Number of standard-normal tensors the reference would have drawn for samples [0, lo) (or [lo, hi) when hi is given)."""
    ode = module.gru_ode
    cam = camera_timestamp.detach().to("cpu", torch.float64).tolist()
    lid = lidar_timestamp.detach().to("cpu", torch.float64).tolist() if lidar_present else None
    tgt = target_timestamp.detach().to("cpu", torch.float64).tolist()
    obs_dtype = "float32" if (camera_timestamp.dtype == torch.float32 and (lid is None or lidar_timestamp.dtype == torch.float32)) else "float64"
    tgt_dtype = "float32" if target_timestamp.dtype == torch.float32 else "float64"
    n = 0
    for b in range(lo, hi) if hi is not None else range(lo):
        order = merge_observations(cam[b], None if lid is None else lid[b])
        n += plan_sample([t for t, _, _ in order], tgt[b], module.delta_t, ode.use_variable_ode_step, ode.solver, obs_dtype, tgt_dtype).n_noise
    return n


def sharded_forward(module, future_prediction_input, camera_states, lidar_states, camera_timestamp, lidar_timestamp,
                    target_timestamp, group: Optional[dist.ProcessGroup] = None, gather: bool = True):
    """FuturePredictionODE.forward on this rank's slice of the batch.  Every rank passes the FULL batch tensors' metadata
    (timestamps) and at least its own slice of the states; returns the gathered [B, ...] output (gather=True) or the local
    slice.  No collective is issued before the optional final all_gather."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = camera_timestamp.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    sl = slice(lo, hi)
    module.gru_ode.noise_skip = noise_draws_before(module, camera_timestamp, lidar_timestamp, target_timestamp,
                                                   lidar_states is not None, lo)
    x, aux = module(future_prediction_input[sl], camera_states[sl], None if lidar_states is None else lidar_states[sl],
                    camera_timestamp[sl], None if lidar_timestamp is None else lidar_timestamp[sl], target_timestamp[sl])
    # leave the generator where the single-process run would leave it: the draws of the samples AFTER this shard are skipped
    # too, so that every rank has consumed the whole batch's stream and the next call starts from the same point everywhere
    after = noise_draws_before(module, camera_timestamp, lidar_timestamp, target_timestamp, lidar_states is not None, hi, B)
    after += module.gru_ode.noise_skip          # an empty shard never drew: its pending skip is still unconsumed
    module.gru_ode.noise_skip = 0
    if after:
        module.gru_ode.discard_noise(after, camera_states)
    if not gather or world == 1:
        return x, aux
    sizes = [shard_bounds(B, r, world) for r in range(world)]
    parts = [torch.empty((b - a,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device) for a, b in sizes]
    dist.all_gather(parts, x.contiguous(), group=group) if len({b - a for a, b in sizes}) == 1 else _all_gather_ragged(parts, x, group)
    return torch.cat(parts, dim=0), aux


def _all_gather_ragged(parts, x, group):
    for r, buf in enumerate(parts):
        if r == (dist.get_rank(group)):
            buf.copy_(x)
        dist.broadcast(buf, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
