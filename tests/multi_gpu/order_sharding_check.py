"""One target composed from draw-order slices on N GPUs (BASELINE config 5, SURVEY 8e batch sharding into one target). Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi_gpu/order_sharding_check.py
Rank r tessellates and submits only its contiguous slice of the scene's draw order; tiles are handed from rank to rank and the
last rank of each tile's chain stores the finished tile everywhere. Every rank then checks that ITS copy of the frame is
bit-identical to the frame the CPU ORACLE produces for the WHOLE scene (rank 0 runs the oracle, digests are broadcast) and to
the frame one GPU renders alone. Scenes: dashed strokes in grid cells (config 5's generator: neighbouring slices overlap along
the cell borders) and overlapping translucent fills in random order (every tile is touched by several ranks).
Prints one JSON line from rank 0; exit code 0 = identical."""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from contrast_renderer_b200 import renderer as R, scenes, sharding  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def main() -> int:
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n_paths = int(os.environ.get("CR_ORDER_CHECK_PATHS", "40000"))
    strokes = scenes.dashed_rational_strokes(n_paths, extent=(1920, 1080), paths_per_shape=500)
    fills = scenes.mixed_fills(600, extent=(1920, 1080), size=(40.0, 260.0), rational=True, seed=31)
    fills.colors[:, 3] = np.linspace(0.3, 1.0, fills.n_shapes, dtype=np.float32)   # translucent: the order of the overs matters
    report = {"check": "order_sharded_target", "n_gpus": world, "scenes": {}}
    ok = True
    for name, scene in (("dashed_strokes", strokes), ("overlapping_fills", fills)):
        config = R.Configuration(device=local)
        reference = [None]
        if rank == 0:
            from oracle import oracle
            oracle.build()
            refs = [oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
                    for i in range(scene.n_shapes)]
            cmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in scenes.stencil_cover_commands(scene.n_shapes)]
            c, s, _, cov = oracle.render(config.to_c(), scene.width, scene.height, refs, cmds, scene.transforms(), scene.colors, threads=oracle.max_threads())
            reference[0] = (digest(c), digest(s), int(cov))
        dist.broadcast_object_list(reference, src=0)

        def render(sharded: bool):
            part = sharding.shard_scene(scene, world, rank) if sharded else scene
            rnd = R.Renderer(config)
            rnd.resize_internal_buffers(scene.width, scene.height)
            stream = torch.cuda.Stream()
            rnd.set_stream(stream.cuda_stream)
            target = sharding.OrderShardedTarget(rnd, stream=stream) if sharded else None
            batch = R.ShapeBatch(rnd, part.dynamic_stroke_options, part.paths, part.shape_path_begin)
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = None
            for _ in range(3):
                torch.cuda.synchronize()
                dist.barrier()
                rp = target.begin_render_pass() if sharded else rnd.begin_render_pass()
                rp.set_instances(part.transforms(), part.colors)
                rp.render_batch(batch, scenes.stencil_cover_commands(part.n_shapes))
                torch.cuda.synchronize()
                dist.barrier()
                start.record(stream)
                if sharded:
                    target.submit(rp)
                else:
                    rp.submit()
                stop.record(stream)
                rnd.synchronize()
                torch.cuda.synchronize()
                ms = start.elapsed_time(stop)
                best = ms if best is None else min(best, ms)
            frame = (rnd.read_color(), rnd.read_stencil(), int(rnd.stats().covered_samples))
            if sharded:
                target.close()
            batch.close()
            rnd.close()
            return frame, best

        (c1, s1, cov1), ms_single = render(False)
        (cn, sn, covn), ms_sharded = render(True)
        same = digest(cn) == reference[0][0] and digest(sn) == reference[0][1] and np.array_equal(c1.view(np.uint32), cn.view(np.uint32)) and np.array_equal(s1, sn)
        t = torch.tensor([int(same), covn], dtype=torch.int64, device=f"cuda:{local}")
        dist.all_reduce(t)
        ms = torch.tensor([ms_sharded], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        identical = int(t[0].item()) == world
        covered_ok = int(t[1].item()) == cov1 == reference[0][2]
        ok = ok and identical and covered_ok
        report["scenes"][name] = {"identical_on_every_rank": identical, "checked_against": "CPU oracle frame of the WHOLE scene (sha256) and the single-GPU frame",
                                  "covered_samples_oracle": reference[0][2], "covered_samples_sum_over_ranks": int(t[1].item()),
                                  "submit_ms_single_gpu": round(ms_single, 3), "submit_ms_sharded_max_over_ranks": round(float(ms.item()), 3),
                                  "paths": scene.paths.n_paths, "shapes": scene.n_shapes}
    report["ok"] = ok
    if rank == 0:
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
