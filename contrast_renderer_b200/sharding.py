"""Multi-GPU plumbing of the path: one process per GPU, path instances sharded by batch (north_star).

Shapes are independent except for draw order into a shared target, so a batch of Shapes is cut into contiguous draw-order
slices, one per rank; each rank tessellates and rasterises its slice with its own renderer into its own target. No
collective is needed on the data path. torch.distributed (NCCL on GPUs, gloo in the CPU tests) only carries the barrier and
the max-over-ranks / sum-over-ranks reductions of the measurements.
"""
from __future__ import annotations

import dataclasses
from typing import Tuple

import numpy as np

from .path import PathSoA
from .scenes import Scene


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of n_items for `rank` of `world` (the first n % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def slice_paths(soa: PathSoA, lo: int, hi: int) -> PathSoA:
    """The paths [lo, hi) of a PathSoA as a self-contained PathSoA (cursor tables rebased)."""
    seg_lo, seg_hi = int(soa.segment_begin[lo]), int(soa.segment_begin[hi])
    segments = []
    type_begin = np.zeros((5, hi - lo + 1), np.uint32)
    for t in range(5):
        t_lo, t_hi = int(soa.type_begin[t, lo]), int(soa.type_begin[t, hi])
        segments.append(np.ascontiguousarray(soa.segments[t][t_lo:t_hi]))
        type_begin[t] = soa.type_begin[t, lo:hi + 1] - t_lo
    return PathSoA(np.ascontiguousarray(soa.start[lo:hi]), (soa.segment_begin[lo:hi + 1] - seg_lo).astype(np.uint32),
                   np.ascontiguousarray(soa.segment_types[seg_lo:seg_hi]), type_begin, segments, np.ascontiguousarray(soa.stroke_options[lo:hi]))


def shard_scene(scene: Scene, world: int, rank: int) -> Scene:
    """The contiguous draw-order slice of `scene`'s Shapes that `rank` owns."""
    s_lo, s_hi = shard_range(scene.n_shapes, world, rank)
    p_lo, p_hi = int(scene.shape_path_begin[s_lo]), int(scene.shape_path_begin[s_hi])
    return dataclasses.replace(
        scene, paths=slice_paths(scene.paths, p_lo, p_hi), shape_path_begin=(scene.shape_path_begin[s_lo:s_hi + 1] - p_lo).astype(np.uint32),
        colors=None if scene.colors is None else np.ascontiguousarray(scene.colors[s_lo:s_hi]),
        origins=None if scene.origins is None else np.ascontiguousarray(scene.origins[s_lo:s_hi]))


def reduce_measurement(elapsed_ms: float, paths: int, covered: int, device=None) -> Tuple[float, int, int]:
    """(max over ranks of elapsed_ms, sum of paths, sum of covered samples); identity without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elapsed_ms), int(paths), int(covered)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s = torch.tensor([paths, covered], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t.item()), int(s[0].item()), int(s[1].item())
