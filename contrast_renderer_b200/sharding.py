"""Multi-GPU plumbing of the path: one process per GPU (north_star). Two ways to shard:

* BATCH sharding (independent path instances, bench.py --gpus N): a batch of Shapes is cut into contiguous draw-order
  slices, one per rank; each rank tessellates and rasterises its slice with its own renderer into its own target. No
  collective on the data path.
* TILE sharding (one render target spanning the box, BASELINE config 4): every rank holds the (small) geometry and records
  the same pass; tile (tx, ty) is owned by rank (tx + ty) % world; a rank bins and rasterises its own tiles only and the
  tile kernel stores every finished tile into all ranks' attachments with P2P stores over NVLink (CUDA IPC mappings of the
  peers' attachments), so the "gather" is fused into the raster kernel's epilogue. `TileShardedTarget` does the handle
  exchange and the two stream-ordered barriers per pass.

torch.distributed (NCCL on GPUs, gloo in the CPU tests) carries the handle exchange, the barriers and the max-over-ranks /
sum-over-ranks reductions of the measurements.
"""
from __future__ import annotations

import dataclasses
from typing import Tuple

import numpy as np

from . import _abi
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


# ------------------------------------------------------------------------------------------------ tile sharding
def tile_owner(tx, ty, world: int):
    """Owner rank of tile (tx, ty): the same function as cr_tile_owned (csrc/raster.h). Diagonal interleave: every rank
    gets every world-th tile of every tile row and column, which balances any spatially coherent scene."""
    return (np.asarray(tx) + np.asarray(ty)) % world


def owned_tile_mask(width: int, height: int, world: int, rank: int, tile: int = 16) -> np.ndarray:
    """[tiles_y, tiles_x] bool: the tiles `rank` rasterises."""
    tiles_x, tiles_y = (width + tile - 1) // tile, (height + tile - 1) // tile
    ty, tx = np.mgrid[0:tiles_y, 0:tiles_x]
    return tile_owner(tx, ty, world) == rank


def exchange_handles(local: bytes, group=None):
    """all_gather of the ranks' IPC handle blobs; returns the list indexed by rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, local, group=group)
    return out


class TileShardedTarget:
    """One render target spanning the ranks of the default process group. Usage, identically on every rank:

        target = TileShardedTarget(renderer)            # after renderer.resize_internal_buffers(...)
        rp = target.begin_render_pass()                 # clears, then barrier: nobody writes into a target still being cleared
        ... record the same draws on every rank ...
        target.submit(rp)                               # own tiles -> all ranks, then barrier: the frame is complete everywhere

    The renderer is moved onto `self.stream`, a torch CUDA stream of its own (the legacy default stream has handle 0,
    which cr_renderer_set_stream reads as "use your own stream"), and the NCCL barriers are issued with that stream
    current, so they are ordered with the renderer's kernels on the device; the host never blocks.
    """

    def __init__(self, renderer, group=None, stream=None):
        import torch
        import torch.distributed as dist
        self.renderer, self.group = renderer, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        device = torch.device(f"cuda:{renderer.config.device}")
        self.stream = stream if stream is not None else torch.cuda.Stream(device)
        assert self.stream.cuda_stream != 0, "the legacy default stream cannot be shared with the renderer"
        self._token = torch.zeros(1, device=device)
        renderer.set_stream(self.stream.cuda_stream)
        renderer.set_tile_sharding(self.world, self.rank)
        handles = exchange_handles(renderer.export_attachments(), group)
        for peer, blob in enumerate(handles):
            if peer != self.rank:
                renderer.import_peer_attachments(peer, blob)
        self.barrier()

    def barrier(self) -> None:
        import torch
        import torch.distributed as dist
        with torch.cuda.stream(self.stream):
            dist.all_reduce(self._token, group=self.group)   # stream-ordered: completes once every rank's earlier work is done

    def begin_render_pass(self, clear_color: bool = True, clear_stencil: bool = True):
        rp = self.renderer.begin_render_pass(clear_color, clear_stencil)
        self.barrier()
        return rp

    def submit(self, render_pass) -> None:
        render_pass.submit()
        self.barrier()

    def close(self) -> None:
        self.barrier()
        self.renderer.synchronize()
        self.renderer.set_tile_sharding(1, 0)


# ----------------------------------------------------------------------------------------------- order sharding
class OrderShardedTarget:
    """ONE render target composed from draw-order slices over the ranks of the default process group (BASELINE config 5:
    "path instances shard across GPUs by batch"). Rank r tessellates and submits only `shard_scene(scene, world, r)`;
    per tile the ranks whose slices touch it hand the tile state from rank to rank over NVLink (the tile kernel waits for its
    predecessor, rasterises on top, stores into its successor's attachments) and the last one stores the finished tile into
    every rank's attachments: the composed frame is what a single GPU produces for the whole scene, bit for bit. Usage,
    identically on every rank:

        target = OrderShardedTarget(renderer)           # after renderer.resize_internal_buffers(...)
        rp = target.begin_render_pass()                 # barrier: nobody still reads the previous frame
        ... record THIS rank's slice ...
        target.submit(rp)                               # barrier: the frame is complete everywhere
    """

    def __init__(self, renderer, group=None, stream=None):
        import torch
        import torch.distributed as dist
        self.renderer, self.group = renderer, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        device = torch.device(f"cuda:{renderer.config.device}")
        self.stream = stream if stream is not None else torch.cuda.Stream(device)
        assert self.stream.cuda_stream != 0, "the legacy default stream cannot be shared with the renderer"
        self._token = torch.zeros(1, device=device)
        renderer.set_stream(self.stream.cuda_stream)
        renderer.set_order_sharding(self.world, self.rank)
        handles = exchange_handles(renderer.export_attachments() + renderer.export_exchange(), group)
        n = _abi.CR_IPC_HANDLE_BYTES
        for peer, blob in enumerate(handles):
            if peer != self.rank:
                renderer.import_peer_attachments(peer, blob[:2 * n])
                renderer.import_peer_exchange(peer, blob[2 * n:])
        self.barrier()

    def barrier(self) -> None:
        import torch
        import torch.distributed as dist
        with torch.cuda.stream(self.stream):
            dist.all_reduce(self._token, group=self.group)   # stream-ordered: completes once every rank's earlier work is done

    def begin_render_pass(self, clear_color: bool = True, clear_stencil: bool = True):
        rp = self.renderer.begin_render_pass(clear_color, clear_stencil)
        self.barrier()
        return rp

    def submit(self, render_pass) -> None:
        render_pass.submit()
        self.barrier()

    def close(self) -> None:
        self.barrier()
        self.renderer.synchronize()
        self.renderer.set_order_sharding(1, 0)


def tile_chains(touched: np.ndarray):
    """The hand-off chains of an order-sharded target as the tile kernel derives them. `touched`: [world, n_tiles] bool.
    Returns (predecessor, successor) int arrays [world, n_tiles]: the rank a rank receives the tile from / hands it to, -1 where
    there is none (or where the rank does not touch the tile)."""
    world, n_tiles = touched.shape
    pred = np.full((world, n_tiles), -1, np.int64)
    succ = np.full((world, n_tiles), -1, np.int64)
    last = np.full(n_tiles, -1, np.int64)
    for r in range(world):
        pred[r] = np.where(touched[r], last, -1)
        last = np.where(touched[r], r, last)
    nxt = np.full(n_tiles, -1, np.int64)
    for r in range(world - 1, -1, -1):
        succ[r] = np.where(touched[r], nxt, -1)
        nxt = np.where(touched[r], r, nxt)
    return pred, succ
