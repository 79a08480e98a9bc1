"""Tile-sharded single target on N GPUs (BASELINE config 4). Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi_gpu/tile_sharding_check.py
Every rank renders the same scripted scene into ONE target spanning the ranks (own tiles only, finished tiles stored into
all ranks' attachments by the tile kernel) and then checks that ITS copy of the complete frame is bit-identical to the
frame the CPU ORACLE produces for the same scene (rank 0 runs the oracle and broadcasts the frame digests) and to the frame
a single-GPU renderer produces. Prints one JSON line from rank 0; exit code 0 = identical."""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from contrast_renderer_b200 import renderer as R, scenes, sharding  # noqa: E402


def main() -> int:
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n_instances = int(os.environ.get("CR_TILE_CHECK_INSTANCES", "200"))
    extent = (1920, 1080)
    scene = scenes.tiger_like(n_instances, extent=extent, instance_px=(60.0, 260.0))
    config = R.Configuration(device=local, alpha_layer_count=2)

    kernel_ms = {}
    cleared = [False]

    def render(sharded: bool, repeats: int = 1):
        rnd = R.Renderer(config)
        rnd.resize_internal_buffers(scene.width, scene.height)
        stream = torch.cuda.Stream()                   # the renderer's work and the CUDA events below go to this stream
        rnd.set_stream(stream.cuda_stream)
        target = sharding.TileShardedTarget(rnd, stream=stream) if sharded else None
        batch = R.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
        rnd.enable_timing(True)
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = []
        for _ in range(repeats):
            torch.cuda.synchronize()
            if sharded:
                dist.barrier()
            rp = target.begin_render_pass() if sharded else rnd.begin_render_pass()
            rp.set_instances(scene.transforms, scene.colors)
            scene.record(rp, batch)       # host-side recording (12 k draws through ctypes) is not what is being compared
            torch.cuda.synchronize()
            dist.barrier()                # the ranks finish recording at different times: start the clocks together
            start.record(stream)
            if sharded:
                target.submit(rp)
            else:
                rp.submit()
            stop.record(stream)
            rnd.synchronize()
            torch.cuda.synchronize()
            ms.append(start.elapsed_time(stop))
        st = rnd.stats()
        frame = (rnd.read_color(), rnd.read_stencil(), int(st.covered_samples), int(st.tile_pairs))
        kernel_ms[sharded] = (round(float(st.last_bin_ms), 3), round(float(st.last_raster_ms), 3))
        if sharded:
            # an empty pass with LoadOp::Clear: every rank may only clear the tiles it owns, in all ranks' attachments
            rp = target.begin_render_pass()
            target.submit(rp)
            rnd.synchronize()
            cleared[0] = bool(not rnd.read_color().any() and not rnd.read_stencil().any())
            target.close()
        batch.close()
        rnd.close()
        return frame, min(ms)

    def digest(a):
        return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()

    # the checker: the CPU oracle's frame of the same scene (test infrastructure; rank 0 only, digests broadcast)
    reference = [None]
    if rank == 0:
        from oracle import oracle
        oracle.build()
        refs = [oracle.shape_from_paths([], scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1])) for i in range(scene.n_shapes)]
        ref_color, ref_stencil, _, ref_covered = oracle.render(config.to_c(), scene.width, scene.height, refs, scene.oracle_commands(), scene.transforms,
                                                               scene.colors, threads=oracle.max_threads())
        reference[0] = (digest(ref_color), digest(ref_stencil), int(ref_covered))
    dist.broadcast_object_list(reference, src=0)
    ref_color_digest, ref_stencil_digest, ref_covered = reference[0]

    (color1, stencil1, covered1, pairs1), ms_single = render(False, 3)
    (colorN, stencilN, coveredN, pairsN), ms_sharded = render(True, 3)
    same_as_oracle = bool(digest(colorN) == ref_color_digest and digest(stencilN) == ref_stencil_digest)
    same = bool(same_as_oracle and np.array_equal(color1.view(np.uint32), colorN.view(np.uint32)) and np.array_equal(stencil1, stencilN))
    stats = torch.tensor([coveredN, pairsN, int(same and cleared[0])], dtype=torch.int64, device=f"cuda:{local}")
    total = stats.clone()
    dist.all_reduce(total)
    t = torch.tensor([ms_sharded], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = int(total[2].item()) == world and int(total[0].item()) == covered1 == ref_covered and int(total[1].item()) == pairs1
    if rank == 0:
        print(json.dumps({"check": "tile_sharded_target", "n_gpus": world, "identical_and_empty_pass_cleared_on_every_rank": int(total[2].item()) == world,
                          "identical": int(total[2].item()) == world, "checked_against": "CPU oracle frame (sha256 of colour and stencil) and the single-GPU frame",
                          "covered_samples_oracle": ref_covered,
                          "covered_samples_single": covered1, "covered_samples_sum_over_ranks": int(total[0].item()),
                          "tile_pairs_single": pairs1, "tile_pairs_sum_over_ranks": int(total[1].item()),
                          "submit_ms_single_gpu": round(ms_single, 3), "submit_ms_sharded_max_over_ranks": round(float(t.item()), 3),
                          "bin_raster_ms_single": kernel_ms[False], "bin_raster_ms_sharded_rank0": kernel_ms[True],
                          "scene": f"tiger_like({n_instances}) {extent[0]}x{extent[1]}, {len(scene.script)} draws", "ok": ok}))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
