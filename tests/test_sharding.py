"""The N > 1 path on CPU: world_size-2 gloo process group, batch sharding of one scene, measurement reductions."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from contrast_renderer_b200 import scenes, sharding


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 317, 1000):
        for world in (1, 2, 3, 8):
            slices = [sharding.shard_range(n, world, r) for r in range(world)]
            assert slices[0][0] == 0 and slices[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(slices, slices[1:]))
            sizes = [hi - lo for lo, hi in slices]
            assert max(sizes) - min(sizes) <= 1


def test_sliced_paths_tessellate_like_the_whole(oracle):
    """A rank's slice of the scene gives, Shape by Shape, the same vertex / index bytes as the unsharded scene."""
    scene = scenes.mixed_fills(60, rational=True, paths_per_shape=4, seed=8)
    whole = [oracle.shape_from_paths([], scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1])) for i in range(scene.n_shapes)]
    seen = 0
    for rank in range(3):
        part = sharding.shard_scene(scene, 3, rank)
        lo, hi = sharding.shard_range(scene.n_shapes, 3, rank)
        assert part.n_shapes == hi - lo and np.array_equal(part.colors, scene.colors[lo:hi])
        for i in range(part.n_shapes):
            got = oracle.shape_from_paths([], part.paths, int(part.shape_path_begin[i]), int(part.shape_path_begin[i + 1]))
            assert np.array_equal(got.vertex_buffer, whole[lo + i].vertex_buffer) and np.array_equal(got.index_buffer, whole[lo + i].index_buffer)
            seen += 1
    assert seen == scene.n_shapes


def _worker(rank: int, world: int, port: int, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        scene = scenes.glyph_like_fills(600, extent=(640, 240), glyphs_per_shape=50)
        part = sharding.shard_scene(scene, world, rank)
        # every rank tessellates its own slice (the CPU stand-in for its renderer) and reports its own clock
        nbytes, _ = oracle.tessellate_batch([], part.paths, part.shape_path_begin, threads=1)
        ms, paths, covered = sharding.reduce_measurement(10.0 * (rank + 1), part.paths.n_paths, nbytes)
        gathered = [None] * world
        dist.all_gather_object(gathered, (part.n_shapes, part.paths.n_paths))
        dist.barrier()
        if rank == 0:
            out.put((ms, paths, covered, gathered, scene.n_shapes, scene.paths.n_paths))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_reduction(built):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ms, paths, nbytes, gathered, n_shapes, n_paths = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 20.0                                     # max over ranks
    assert paths == n_paths                               # every path is owned by exactly one rank
    assert sum(g[0] for g in gathered) == n_shapes and sum(g[1] for g in gathered) == n_paths
    assert nbytes > 0


# ------------------------------------------------------------------------------------------------ tile sharding (config 4)
def test_tile_ownership_is_a_balanced_partition():
    for width, height in ((3840, 2160), (640, 400), (17, 33)):
        for world in (1, 2, 3, 4, 8):
            masks = [sharding.owned_tile_mask(width, height, world, r) for r in range(world)]
            total = np.sum(masks, axis=0)
            assert (total == 1).all(), "every tile has exactly one owner"
            counts = [int(m.sum()) for m in masks]
            assert max(counts) - min(counts) <= max(masks[0].shape), "diagonal interleave is balanced to within one tile row"
            if world > 1 and masks[0].shape[1] >= world:
                assert all(m[0, :world].sum() == 1 for m in masks), "every rank owns one of any `world` consecutive tiles of a row"


def _tile_worker(rank: int, world: int, port: int, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        from contrast_renderer_b200 import renderer as R
        # the CPU stand-in for a rank: the oracle renders the replicated scene, the rank keeps the tiles it owns and
        # "stores" them into every rank's frame (here: an all_gather of the owned tiles), like K3's peer stores
        scene = scenes.tiger_like(3, extent=(160, 96), instance_px=(60.0, 120.0))
        refs = [oracle.shape_from_paths([], scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1])) for i in range(scene.n_shapes)]
        config = R.Configuration(alpha_layer_count=2)
        color, stencil, _, _ = oracle.render(config.to_c(), scene.width, scene.height, refs, scene.oracle_commands(), scene.transforms, scene.colors)
        mask = np.kron(sharding.owned_tile_mask(scene.width, scene.height, world, rank), np.ones((16, 16), bool))[:scene.height, :scene.width]
        mine = np.where(mask[..., None, None], color.reshape(scene.height, scene.width, 1, 4), 0.0)
        handles = sharding.exchange_handles(bytes([rank]) * 128)
        parts = [None] * world
        dist.all_gather_object(parts, (mine, mask))
        frame = np.zeros_like(mine)
        for part, m in parts:
            frame[m] = part[m]
        if rank == 0:
            out.put((handles, np.array_equal(frame.reshape(color.shape), color), float(np.abs(color).sum())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_tile_sharded_frame(built):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_tile_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    handles, same, energy = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert handles == [bytes([0]) * 128, bytes([1]) * 128]   # blobs arrive indexed by rank
    assert same and energy > 0, "the owned tiles of the two ranks compose the whole frame"


def test_tile_chains_of_an_order_sharded_target():
    """Host-side restatement of the hand-off chains the tile kernel derives from the touched-tile bitmaps (raster.cu): per tile the
    touching ranks in rank order; every touched tile has exactly one chain head (no predecessor) and one tail (no successor)."""
    from contrast_renderer_b200 import sharding
    rng = np.random.default_rng(4)
    touched = rng.random((5, 300)) < 0.4
    pred, succ = sharding.tile_chains(touched)
    for t in range(touched.shape[1]):
        ranks = [r for r in range(5) if touched[r, t]]
        for i, r in enumerate(ranks):
            assert pred[r, t] == (ranks[i - 1] if i else -1) and succ[r, t] == (ranks[i + 1] if i + 1 < len(ranks) else -1)
        for r in range(5):
            if not touched[r, t]:
                assert pred[r, t] == -1 and succ[r, t] == -1
    heads = ((pred == -1) & touched).sum(0)
    tails = ((succ == -1) & touched).sum(0)
    assert np.array_equal(heads, touched.any(0).astype(int)) and np.array_equal(tails, touched.any(0).astype(int))


def test_draw_order_slices_cover_the_scene():
    """shard_scene: the ranks' slices are contiguous, disjoint and complete in draw order (what order sharding relies on)."""
    from contrast_renderer_b200 import scenes, sharding
    scene = scenes.dashed_rational_strokes(5000, extent=(640, 360), paths_per_shape=250)
    world = 3
    parts = [sharding.shard_scene(scene, world, r) for r in range(world)]
    assert sum(p.n_shapes for p in parts) == scene.n_shapes and sum(p.paths.n_paths for p in parts) == scene.paths.n_paths
    assert np.array_equal(np.concatenate([p.colors for p in parts]), scene.colors)
    assert np.array_equal(np.concatenate([p.paths.start for p in parts]), scene.paths.start)
    assert np.array_equal(np.concatenate([p.transforms() for p in parts]), scene.transforms())
