"""Multi-GPU model of the NCC tracking path: one process per GPU, independent tracks sharded by stream.

The path has no exchange step: a track owns {frames, bbox, template} and only depends on its own
previous frame, so ranks never communicate on the data path.  Track i lives on rank i mod world; each rank drives its
own pvt_ctx (the C ABI is per device).  The only collective is an OPTIONAL final gather of the per-frame records
(tracks x frames x 32 B) over torch.distributed -- NCCL on GPUs, gloo in the CPU tests -- off the timed path.
Nothing here computes NCC: `step_fn` is whatever runs one time step for the local tracks (Tracker.step on a GPU).
"""
from __future__ import annotations

import numpy as np


def shard_tracks(n_tracks: int, world: int, rank: int) -> list[int]:
    """Global track ids owned by `rank`: i -> rank i mod world (round-robin keeps shards within one track of each other)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return list(range(rank, n_tracks, world))


def owner_of(track: int, world: int) -> tuple[int, int]:
    """(rank, local slot) of a global track id."""
    return track % world, track // world


def local_count(n_tracks: int, world: int, rank: int) -> int:
    return len(range(rank, n_tracks, world))


def gather_records(local: np.ndarray, n_tracks: int, world: int, rank: int, dist=None) -> np.ndarray:
    """All-gather per-track record arrays.

    local: [n_local, ...] records of this rank's tracks in shard order (any numeric / structured dtype).
    Returns [n_tracks, ...] in GLOBAL track order on every rank.  dist: torch.distributed (None or world==1: no comm).
    Shards may differ by one track, so every rank pads to the largest shard before the fixed-size all_gather.
    """
    n_local = local_count(n_tracks, world, rank)
    if local.shape[0] != n_local:
        raise ValueError(f"rank {rank} holds {local.shape[0]} records, expected {n_local}")
    if world == 1 or dist is None:
        return local.copy()
    import torch

    n_max = local_count(n_tracks, world, 0)
    raw = np.zeros((n_max,) + local.shape[1:], local.dtype)
    raw[:n_local] = local
    flat = torch.from_numpy(raw.view(np.uint8).reshape(-1).copy())
    if dist.get_backend() == "nccl":
        flat = flat.cuda()
    outs = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(outs, flat)
    full = np.zeros((n_tracks,) + local.shape[1:], local.dtype)
    for r, o in enumerate(outs):
        part = o.cpu().numpy().view(local.dtype).reshape((n_max,) + local.shape[1:])
        ids = shard_tracks(n_tracks, world, r)
        full[ids] = part[:len(ids)]
    return full


def run_sharded(n_tracks, n_steps, world, rank, make_local, step_fn, dist=None):
    """Drive `n_steps` time steps for this rank's shard and gather every track's records at the end.

    make_local(track_ids) -> state ; step_fn(state, k) -> records [n_local, ...] for time step k.
    Returns [n_steps, n_tracks, ...] in global order (one final gather, nothing exchanged per step).
    """
    ids = shard_tracks(n_tracks, world, rank)
    state = make_local(ids)
    per_step = [np.asarray(step_fn(state, k)) for k in range(n_steps)]
    local = np.stack(per_step, 1) if per_step else np.zeros((len(ids), 0))
    full = gather_records(local, n_tracks, world, rank, dist)
    return np.swapaxes(full, 0, 1)
