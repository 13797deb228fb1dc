"""Multi-GPU plumbing for the "batches of independent sequences, one scene per GPU" mode.

The fusion path shards by scene: every rank owns whole scenes (hash table + voxel blocks + tracking state stay on one
GPU), so there is no collective on the data path.  torch.distributed is used only to agree on the partition, to
barrier around the timed region and to combine the per-rank timings (max over ranks).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def partition_scenes(n_scenes: int, world_size: int, rank: int):
    """Contiguous, balanced assignment of scene ids to ranks (first ranks get the remainder)."""
    base, rem = divmod(n_scenes, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def sequence_start_for_scene(scene_id: int) -> int:
    """BASELINE configs[3]: 64 copies of the sequence with phase-shifted trajectories."""
    return (7 * scene_id) % 100


def combine_timing(local_ms: float, local_frames: int, device=None):
    """Returns (max over ranks of the elapsed ms, total frames over all ranks)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(local_ms), int(local_frames)
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    n = torch.tensor([int(local_frames)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item())


def aggregate_frames_per_second(local_ms: float, local_frames: int, device=None) -> float:
    ms, frames = combine_timing(local_ms, local_frames, device)
    return frames / (ms * 1e-3)
