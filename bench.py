#!/usr/bin/env python
"""bench.py — paths/sec and covered-Mpixel/s of the tessellate -> stencil-then-cover hot path, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--impl reference]

--config picks the BASELINE.json configuration (default 3, the 100k-glyph 4K scene the metric is quoted on):
  1  1k closed cubic paths -> stroke tessellation to vertex buffers only          (a step = Shape::from_paths)
  2  10k mixed line / quadratic / cubic fills, non-zero winding, 1920x1080
  3  100k TTF glyph instances through the text front-end, 3840x2160
  4  1000 placed copies of the 240-path group, nested clips + opacity groups, 3840x2160
  5  1M dashed round-joined rational-cubic strokes, 7680x4320
A "step" is one whole pass of the hot path over the scene: Shape::from_paths for every shape of the scene (one batched
launch sequence; config 4 re-tessellates its 24-Shape group) followed by one render pass into a cleared target.
  value            : paths/s with the path arrays already resident in HBM when the timed region starts; K steps back to back,
                     nothing read back in between (CUDA events on the renderer's stream)
  e2e              : the same through the C-ABI with HOST (pinned) input arrays — host->device staging of every input inside the
                     timed region — and a device->host read of the pass result (the counters) every step
  e2e_with_frame   : e2e plus the device->host copy of the colour attachment (RGBA8 target: the frame a presenter consumes)
N > 1: configs 1, 2, 3 shard path instances across GPUs by batch (north_star): each rank owns an independent scene of the same
size (weak scaling, no data-path collective); config 4 is ONE target tile-sharded over the ranks and config 5 ONE target composed
from the ranks' draw-order slices (both strong scaling: K3 stores tiles into the other ranks' attachments over NVLink). At N > 1 the default run also renders a config-4 frame
tile-sharded over the N ranks and checks every rank's copy against the CPU oracle's frame (`tile_sharded_check`).
`--impl reference`: the reference is a Rust crate with no toolchain in this image, so the reference arm is its CPU
restatement (oracle/, `kind: "port"`): tessellation on one thread like the reference's loop, raster on all host threads.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------------ workloads
class Workload:
    """One BASELINE configuration: the scene, the renderer configuration, how a pass is recorded, the oracle's commands."""

    def __init__(self, index: int, rank: int, scale: float = 1.0, world: int = 1):
        from contrast_renderer_b200 import scenes
        self.index = index
        seed_shift = 1000 * rank if index in (1, 2, 3) else 0   # configs 4 and 5 at N > 1: ONE scene, one target
        self.alpha_layers = 0
        self.tess_only = False
        self.scripted = False
        if index == 1:
            n = max(8, int(1000 * scale))
            self.scene = scenes.closed_cubic_strokes(n, seed=scenes.SEED0 + 1 + seed_shift)
            self.tess_only = True
            self.name = f"{n} closed paths of 4 integral cubics, stroked (width 1-8 px, miter, UniformTangentAngle(0.1)): tessellation to vertex / index / hull buffers only"
        elif index == 2:
            n = max(8, int(10000 * scale))
            self.scene = scenes.mixed_fills(n, seed=scenes.SEED0 + 2 + seed_shift)
            self.name = f"{n} mixed line / quadratic / cubic filled paths, non-zero winding (4 bits), 1920x1080, one Shape (Stencil+Color) per path"
        elif index == 3:
            n = max(160, int(100000 * scale))
            self.scene = scenes.text_glyphs(n, seed=scenes.SEED0 + 3 + seed_shift, extent=(3840, 2160), glyphs_per_shape=160)
            self.name = (f"{n} TTF glyph instances via the text front-end (paths_of_text layout, OpenSans outlines), 12 px glyphs, 3840x2160, "
                         "one Shape (Stencil+Color) per 160-glyph run")
        elif index == 4:
            n = max(2, int(1000 * scale))
            self.scene = scenes.tiger_like(n, seed=scenes.SEED0 + 4)   # ONE target for all ranks: the same scene everywhere
            self.alpha_layers = 2
            self.scripted = True
            self.name = (f"{n} placed copies of a 24-Shape / 240-path constructor-built group (rational conics and cubics), 3 nested clips and 2 nested "
                         "opacity groups per copy, 3840x2160")
        elif index == 5:
            n = max(1000, int(1000000 * scale))
            self.scene = scenes.dashed_rational_strokes(n, seed=scenes.SEED0 + 5 + seed_shift)
            self.name = f"{n} dashed stroked open paths of 2 rational cubics, round joins and caps, UniformTangentAngle(0.2), 7680x4320, one Shape per 1000 paths"
        else:
            raise SystemExit(f"--config {index}: BASELINE.json has configurations 1..5")
        self.total_paths = self.scene.paths.n_paths
        if index == 5 and world > 1:   # this rank's contiguous slice of the draw order (the whole scene is composed into one target)
            from contrast_renderer_b200 import sharding
            self.scene = sharding.shard_scene(self.scene, world, rank)
        s = self.scene
        self.width, self.height = s.width, s.height
        self.transforms = s.transforms if self.scripted else s.transforms()
        self.colors = s.colors
        self.commands = None if self.scripted or self.tess_only else scenes.stencil_cover_commands(s.n_shapes)
        self.n_instances = len(self.transforms)
        # path INSTANCES drawn per step: config 4 draws its 240 paths once per placed copy
        self.paths_per_step = self.total_paths * (n if index == 4 else 1)   # of the whole scene (configs 4, 5 at N > 1: counted once, not per rank)

    def configuration(self, cr, device: int, color_format=None):
        kw = dict(device=device, alpha_layer_count=self.alpha_layers)
        if color_format is not None:
            kw["color_format"] = color_format
        return cr.Configuration(**kw)

    def record(self, rp, batch) -> None:
        if self.scripted:
            self.scene.record(rp, batch)
        else:
            rp.render_batch(batch, self.commands)

    def oracle_commands(self):
        if self.scripted:
            return self.scene.oracle_commands()
        return [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in self.commands]

    def config_dict(self):
        """Identical in the b200 and the reference arm."""
        return {"workload": f"BASELINE config {self.index}: {self.name}", "baseline_config": self.index, "paths_per_gpu": self.scene.paths.n_paths,
                "path_instances_per_step_per_gpu": self.paths_per_step, "segments_per_gpu": self.scene.paths.n_segments, "shapes_per_gpu": self.scene.n_shapes,
                "width": self.width, "height": self.height, "msaa_sample_count": 1, "winding_counter_bits": 4, "clip_nesting_counter_bits": 4,
                "color_format": "rgba32f"}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:   # driver-written: STREAM-style copy bandwidth of this pool's B200s (B200_PROFILING.md)
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except (OSError, KeyError, TypeError, ValueError):
        return 6650.0, "fallback"   # the recipe's stated fallback


def measured_traffic(kernel: str, config_index: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed `ncu --set full` capture of this
    command (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None when no capture is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            table = json.load(f)
        return table.get(f"config{config_index}", {}).get(kernel)
    except (OSError, ValueError):
        return None


class ClockSampler:
    """`nvidia-smi -lms 10` in the background. It needs ~0.1 s to come up, longer than a short timed region, so it is started
    before the warm-up and the samples are selected by their timestamps: those taken inside the timed window [mark_begin(),
    mark_end()]; only if the window is too short to hold one are the (identically loaded) warm-up samples used, and the
    result says so in "window"."""
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.t_begin = self.t_end = None

    def mark_begin(self):
        import datetime
        self.t_begin = datetime.datetime.now()

    def mark_end(self):
        import datetime
        self.t_end = datetime.datetime.now()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "10", "-i", str(self.device_index)],
                                         stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        import datetime
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.file.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                stamp = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((stamp, float(parts[1]), float(parts[2]), [n for n, v in zip(names, parts[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.file.name)
        inside = [r for r in rows if self.t_begin is not None and self.t_end is not None and self.t_begin <= r[0] <= self.t_end]
        window = "timed region"
        if not inside:
            inside, window = rows, "warm-up + timed region (the timed region is shorter than one sampling period)"
        if inside:
            out.update(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(max(r[2] for r in inside)),
                       reasons=sorted({n for r in inside for n in r[3]}), samples=len(inside), window=window)
        return out


# -------------------------------------------------------------------------------------------------- CPU (oracle) legs
def oracle_step(work: Workload, threads: int):
    """One step of the CPU restatement: sequential tessellation like the reference's loop (src/renderer.rs:187), raster on
    `threads` host threads. Returns (seconds tessellating, seconds rasterising, covered samples)."""
    from oracle import oracle
    from contrast_renderer_b200.renderer import Configuration
    scene = work.scene
    t0 = time.perf_counter()
    shapes = [oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
              for i in range(scene.n_shapes)]
    t1 = time.perf_counter()
    covered = 0
    if not work.tess_only:
        cfg = Configuration(alpha_layer_count=work.alpha_layers).to_c()
        _, _, _, covered = oracle.render(cfg, work.width, work.height, shapes, work.oracle_commands(), work.transforms, work.colors, threads=threads)
    return t1 - t0, time.perf_counter() - t1, covered


def reference_scale(index: int) -> float:
    """Fraction of the configuration one step of the CPU arm processes, so that 25 steps end within a few minutes."""
    return {1: 1.0, 2: 1.0, 3: 1.0, 4: 0.2, 5: 0.02}[index]


def run_reference(args, rank: int):
    """The reference arm: the CPU restatement of the reference on the box's host cores."""
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    threads = oracle.max_threads()
    scale = args.scale if args.scale is not None else reference_scale(args.config)
    full = Workload(args.config, 0, 1.0 if args.scale is None else args.scale)
    work = full if scale == 1.0 or args.scale is not None else Workload(args.config, 0, scale)
    for _ in range(args.warmup):
        oracle_step(work, threads)
    t0 = time.perf_counter()
    tess = raster = 0.0
    covered = 0
    for _ in range(args.steps):
        a, b, covered = oracle_step(work, threads)
        tess, raster = tess + a, raster + b
    dt = time.perf_counter() - t0
    value = work.paths_per_step * args.steps / dt
    sample = (f"{'the whole configuration' if work is full else f'{scale:g} of the configuration'} per step ({work.scene.paths.n_paths} paths, {work.scene.n_shapes} shapes, "
              f"{work.width}x{work.height}); tessellation on 1 thread {tess / args.steps:.3f} s per step (the reference loop is sequential, src/renderer.rs:187), "
              f"raster on {threads} threads {raster / args.steps:.3f} s per step")
    line = {
        "impl": "reference", "metric": "paths/sec", "value": value, "unit": "paths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": full.config_dict(),
        "covered_mpixel_per_s": covered * args.steps / dt / 1e6,
        "cpu_baseline": {"value": value, "unit": "paths/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(index: int, scale_arg, budget_s: float = 12.0):
    """Oracle (the CPU port of the reference) on a bounded sample of the same workload, about `budget_s` seconds."""
    from oracle import oracle
    threads = oracle.max_threads()
    scale = scale_arg if scale_arg is not None else reference_scale(index)
    work = Workload(index, 0, scale)
    reps, tess, raster = 0, 0.0, 0.0
    while reps < 12 and tess + raster < budget_s:
        a, b, _ = oracle_step(work, threads)
        reps, tess, raster = reps + 1, tess + a, raster + b
    return {"value": reps * work.paths_per_step / (tess + raster), "unit": "paths/s", "cores": threads, "kind": "port",
            "sample": f"{reps} x {scale:g} of the configuration ({work.scene.paths.n_paths} paths, {work.width}x{work.height} target); tessellation on 1 thread "
                      f"{tess / reps:.3f} s per pass (the reference loop is sequential, src/renderer.rs:187), raster on {threads} threads {raster / reps:.3f} s per pass"}


# ------------------------------------------------------------------------------------------------------- GPU legs
def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def tile_sharded_check(cr, dist, torch, rank: int, world: int, local_rank: int):
    """A config-4 frame (200 placed copies, 1920x1080) as ONE target tile-sharded over the ranks: every rank's copy of the
    complete frame against the CPU oracle's frame (rank 0 runs the oracle and broadcasts the digests), with the submit time."""
    from contrast_renderer_b200 import scenes, sharding
    scene = scenes.tiger_like(200, extent=(1920, 1080), instance_px=(60.0, 260.0))
    config = cr.Configuration(device=local_rank, alpha_layer_count=2)
    reference = [None]
    if rank == 0:
        from oracle import oracle
        oracle.build()
        refs = [oracle.shape_from_paths([], scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1])) for i in range(scene.n_shapes)]
        c, s, _, cov = oracle.render(config.to_c(), scene.width, scene.height, refs, scene.oracle_commands(), scene.transforms, scene.colors, threads=oracle.max_threads())
        reference[0] = (digest(c), digest(s), int(cov))
    dist.broadcast_object_list(reference, src=0)
    rnd = cr.Renderer(config)
    rnd.resize_internal_buffers(scene.width, scene.height)
    stream = torch.cuda.Stream()
    rnd.set_stream(stream.cuda_stream)
    target = sharding.TileShardedTarget(rnd, stream=stream)
    batch = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for _ in range(3):
        rp = target.begin_render_pass()
        rp.set_instances(scene.transforms, scene.colors)
        scene.record(rp, batch)
        torch.cuda.synchronize()
        dist.barrier()
        start.record(stream)
        target.submit(rp)
        stop.record(stream)
        rnd.synchronize()
        torch.cuda.synchronize()
        best = start.elapsed_time(stop) if best is None else min(best, start.elapsed_time(stop))
    same = digest(rnd.read_color()) == reference[0][0] and digest(rnd.read_stencil()) == reference[0][1]
    covered = int(rnd.stats().covered_samples)
    t = torch.tensor([int(same), covered], dtype=torch.int64, device=f"cuda:{local_rank}")
    dist.all_reduce(t)
    ms = torch.tensor([best], dtype=torch.float64, device=f"cuda:{local_rank}")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    target.close()
    batch.close()
    rnd.close()
    return {"scene": "BASELINE config 4 at 200 copies, 1920x1080, 12 000 draws, one target tile-sharded over the ranks",
            "identical": int(t[0].item()) == world, "checked_against": "CPU oracle frame (sha256 of colour and stencil), every rank's copy",
            "covered_samples_sum_over_ranks": int(t[1].item()), "covered_samples_oracle": reference[0][2], "submit_ms_max_over_ranks": float(ms.item()), "n_gpus": world}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, help="BASELINE.json configuration 1..5 (default 3: the one the metric is quoted on)")
    ap.add_argument("--scale", type=float, default=None, help="debug only: a fraction of the configuration is not the benchmark workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded-check", action="store_true")
    ap.add_argument("--no-pipelining", action="store_true", help="keep every step on one stream (no overlap of consecutive steps)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from contrast_renderer_b200 import _abi, renderer as R, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    work = Workload(args.config, rank, 1.0 if args.scale is None else args.scale, world)
    scene, soa = work.scene, work.scene.paths
    one_target = args.config in (4, 5) and world > 1   # ONE render target spanning the ranks (4: tile sharding, 5: draw-order slices), else an independent scene per rank

    def make_renderer(color_format=None):
        rnd = R.Renderer(work.configuration(R, local_rank, color_format))
        stream = torch.cuda.Stream(dev)   # not the legacy default stream: its handle is 0, which cr_renderer_set_stream reads as "use your own stream"
        assert stream.cuda_stream != 0
        rnd.set_stream(stream.cuda_stream)
        rnd.resize_internal_buffers(work.width, work.height)
        if not args.no_pipelining:
            rnd.set_pipelining(True)   # the rebuild of frame N + 1 overlaps the raster of frame N (inputs are complete before every call here)
        return rnd, stream

    rnd, stream = make_renderer()
    torch.cuda.set_stream(stream)
    target = None
    if one_target:
        rnd.set_pipelining(False)   # the ranks hand tiles to each other inside the pass: one stream, barriers around every submit
        target = sharding.TileShardedTarget(rnd, stream=stream) if args.config == 4 else sharding.OrderShardedTarget(rnd, stream=stream)

    # device-resident and pinned-host copies of every input array
    host_arrays = soa.arrays()
    if not soa.any_stroked():
        host_arrays[9] = host_arrays[9][:0]   # all paths are filled: cr_path_soa.stroke_options = NULL, nothing to copy
    pinned = [torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).pin_memory() for a in host_arrays]
    resident = [t.to(dev) for t in pinned]
    inst_host = [torch.from_numpy(np.ascontiguousarray(work.transforms, np.float32).reshape(-1).copy()).pin_memory(),
                 torch.from_numpy(np.ascontiguousarray(work.colors, np.float32).reshape(-1).copy()).pin_memory()]
    inst_dev = [t.to(dev) for t in inst_host]
    n_commands = 0 if work.tess_only else (len(scene.script) if work.scripted else len(work.commands))
    h2d_bytes = sum(t.numel() for t in pinned) + 4 * len(scene.shape_path_begin)
    if not work.tess_only:
        h2d_bytes += sum(t.numel() * 4 for t in inst_host) + 32 * n_commands

    # Frame pipelining at the application level: the step's pass is submitted AFTER the next step's from_paths has been called
    # (two batch objects alternate), so that the host's wait inside from_paths (for the sizes of the count pass, behind the
    # input copy) overlaps the previous frame's hulls and raster. Every step still is one from_paths + one pass + one result
    # read; the timed regions end with the last step's pass submitted and waited for (flush).
    software_pipelined = not args.no_pipelining and not one_target and not work.tess_only
    state, pending, parity = {}, {}, {}

    def submit_pass(r, batch, device_resident: bool):
        rp = target.begin_render_pass() if (one_target and r is rnd) else r.begin_render_pass()
        if device_resident:
            rp.set_instances(inst_dev[0].data_ptr(), inst_dev[1].data_ptr(), count=work.n_instances, memory_space=_abi.CR_MEM_DEVICE)
        else:
            rp.set_instances(inst_host[0].data_ptr(), inst_host[1].data_ptr(), count=work.n_instances, memory_space=_abi.CR_MEM_HOST)
        work.record(rp, batch)
        if one_target and r is rnd:
            target.submit(rp)
        else:
            rp.submit()

    def step(r, device_resident: bool, in_order: bool = False):
        if device_resident:
            ptrs, space = [t.data_ptr() if t.numel() else 0 for t in resident], _abi.CR_MEM_DEVICE
        else:
            ptrs, space = [t.data_ptr() if t.numel() else 0 for t in pinned], _abi.CR_MEM_HOST
        deferred = software_pipelined and not in_order
        slot = (r, parity.get(r, 0) if deferred else 0)
        if not deferred:
            flush(r)
        batch = state[slot] = R.ShapeBatch(r, scene.dynamic_stroke_options, soa, scene.shape_path_begin, existing=state.get(slot), memory_space=space, pointers=ptrs)
        if work.tess_only:
            return
        if not deferred:
            submit_pass(r, batch, device_resident)
            return
        parity[r] = 1 - parity.get(r, 0)
        previous = pending.get(r)
        pending[r] = (batch, device_resident)
        if previous is not None:
            submit_pass(r, *previous)

    def flush(r):
        previous = pending.pop(r, None)
        if previous is not None:
            submit_pass(r, *previous)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def timed(r, st, steps: int, device_resident: bool, per_step=None, finish=None, in_order: bool = False):
        """K steps back to back on stream `st`, bracketed by a barrier + synchronize on both sides, CUDA-event time, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            step(r, device_resident, in_order)
            if per_step is not None:
                per_step()
        flush(r)   # software pipelining: the last step's pass
        if finish is not None:
            finish()
        e1.record(st)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- warm-up, then the timed regions
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_warm = time.time()
    for _ in range(max(8 if software_pipelined else 3, args.warmup)):   # software pipelining: 2 batch objects x 3 array sets take their first (cold) build here
        step(rnd, True)
    for _ in range(6):   # host inputs: both staging sets and every array set of both batch objects see a host-input build before the timed regions
        step(rnd, False)
    step(rnd, True)
    while rank == 0 and world == 1 and time.time() - t_warm < 0.5:   # untimed: gives nvidia-smi time to come up, under the benchmark's own load
        step(rnd, True)
    flush(rnd)
    rnd.synchronize()
    torch.cuda.synchronize(dev)
    launches0 = int(rnd.stats().kernel_launches)
    sampler.mark_begin()
    ms_dev = timed(rnd, stream, args.steps, True)
    launches = int(rnd.stats().kernel_launches) - launches0

    results = {}

    def read_result():
        results["stats"] = rnd.stats()   # device->host read of the pass result (counters); synchronises the stream

    def read_settled_result():
        # Frame pipelining: every pass copies its counters to the host (32 B, inside the timed region); submit() of step N + 1
        # waits for step N's, so the host consumes the result of step N while step N + 1 runs. The last step's is waited for
        # (read_result) before the end of the timed region.
        results["stats"] = rnd.settled_pass_stats()

    pipelined_e2e = not args.no_pipelining and not one_target and not work.tess_only
    ms_e2e = timed(rnd, stream, args.steps, False, per_step=read_settled_result if pipelined_e2e else read_result, finish=read_result)
    assert work.tess_only or int(results["stats"].covered_samples) > 0
    # one step at a time (the host waits for each step before starting the next): the latency of a step, no overlap of steps
    ms_serial = timed(rnd, stream, min(args.steps, 10), True, per_step=rnd.synchronize, in_order=True) / min(args.steps, 10)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    d2h_result_bytes = 3 * 8 + 2 * 4   # PassCounters

    # per-kernel times (CUDA events inside the library, read back every step): a separate, untimed-for-throughput run
    rnd.enable_timing(True)
    kernel_ms = {"tess": [], "bin": [], "raster": [], "hull_sort": [], "hull_chain": []}
    for _ in range(5):
        step(rnd, True, in_order=True)
        st = rnd.stats()
        for key, v in (("tess", st.last_tess_ms), ("bin", st.last_bin_ms), ("raster", st.last_raster_ms), ("hull_sort", st.last_hull_sort_ms),
                       ("hull_chain", st.last_hull_chain_ms)):
            kernel_ms[key].append(v)
    rnd.enable_timing(False)
    st = rnd.stats()
    covered = int(st.covered_samples)

    # e2e_with_frame: host inputs in, the finished RGBA8 frame out, every step
    frame_line = None
    if not work.tess_only and not one_target:
        rnd8, stream8 = make_renderer(R.ColorFormat.Rgba8Unorm)
        frame = torch.empty(work.width * work.height * 4, dtype=torch.uint8).pin_memory()

        def read_frame():
            rnd8.read_color_texels(frame.data_ptr(), frame.numel())

        # pipelined: the frame of step N is snapshot behind its pass and copied to the host (three pinned frames rotate) while the
        # next steps are rendered; the host waits for frame N - 2 every step and for the last frames before the timed region ends
        frames = [frame, torch.empty_like(frame).pin_memory(), torch.empty_like(frame).pin_memory()]
        tickets = []

        def read_frame_async():
            if len(tickets) == 3:
                rnd8.wait_readback(tickets.pop(0))
            k = read_frame_async.count = getattr(read_frame_async, "count", 0) + 1
            tickets.append(rnd8.read_color_texels_async(frames[k % 3].data_ptr(), frame.numel()))

        def drain_frames():
            while tickets:
                rnd8.wait_readback(tickets.pop(0))

        torch.cuda.set_stream(stream8)
        pipelined_frames = software_pipelined
        for _ in range(8 if pipelined_frames else 3):
            step(rnd8, False, in_order=not pipelined_frames)
            read_frame_async() if pipelined_frames else read_frame()
        flush(rnd8)
        drain_frames()
        if pipelined_frames:
            ms_frame = timed(rnd8, stream8, args.steps, False, per_step=read_frame_async, finish=lambda: (read_frame_async(), drain_frames()))
        else:
            ms_frame = timed(rnd8, stream8, args.steps, False, per_step=read_frame, in_order=True)
        torch.cuda.set_stream(stream)
        frame_line = {"value": None, "unit": "paths/s", "ms_per_step": ms_frame / args.steps, "h2d_bytes_per_step": int(h2d_bytes),
                      "d2h_bytes_per_step": int(frame.numel()), "color_format": "rgba8unorm",
                      "frame_read": ("every step's RGBA8 frame is snapshot behind its pass and copied to pinned host memory on its own stream while the next "
                                     "steps run; the host waits for each frame two steps later and for the last ones before the timed region ends")
                                    if pipelined_frames else "the host waits for every step's frame before it starts the next step"}
        for key in [k for k in state if k[0] is rnd8]:
            state.pop(key).close()
        rnd8.close()

    total_paths, total_covered = work.paths_per_step, covered
    if world > 1:
        t = torch.tensor([work.paths_per_step, covered], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_covered = int(t[1].item())
        total_paths = work.paths_per_step if one_target else int(t[0].item())   # one target: one scene for all ranks; count it once
        if one_target and args.config == 4:
            total_covered = covered   # tile sharding: every rank's counter... is per owner; the sum over ranks is the frame's
            total_covered = int(t[1].item())

    sharded = None
    if world > 1 and not args.no_sharded_check:
        sharded = tile_sharded_check(R, dist, torch, rank, world, local_rank)

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        mean = lambda xs: float(np.mean(xs)) if xs else 0.0
        k_ms = {k: mean(v) for k, v in kernel_ms.items()}
        layout_bytes = int(st.vertex_bytes)
        fb_bytes = work.width * work.height * (1 + 16)
        raster_alg = layout_bytes + 80 * work.n_instances + fb_bytes          # SURVEY §8d B_rast
        tess_alg = int(st.input_bytes) + layout_bytes                         # SURVEY §8d B_tess
        chain_alg = 8 * int(st.proto_hull_points) + 8 * int(st.hull_vertices)

        def roofline_of(name, alg, ms, formula):
            ach = alg / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak if peak else None,
                    "algorithmic_bytes": alg, "algorithmic_bytes_formula": formula, "kernel_ms": ms,
                    "traffic": measured_traffic(name, args.config) if args.scale is None else None}
        # Roofline of the DOMINANT single kernel of the step, timed live with CUDA events on the renderer's stream: K3
        # raster_tiles_kernel with B_rast of SURVEY 8d, or hull_chain_kernel (convex_hull::andrew's chains: 8 B per proto-hull
        # point read + 8 B per hull vertex written), or — tessellation-only config 1 — the emit pass with B_tess.
        tess_emit_ms = max(k_ms["tess"] - k_ms["hull_sort"] - k_ms["hull_chain"], 0.0)
        rooflines = [roofline_of("hull_chain_kernel", chain_alg, k_ms["hull_chain"], "8 B x proto-hull points + 8 B x hull vertices (latency-bound sequential stack machines)"),
                     roofline_of("tess_count+scan+emit", tess_alg, tess_emit_ms, "B_in + B_out of SURVEY 8d (path input + packed vertex / index output)")]
        if not work.tess_only:
            # timed together: tile_prims_kernel (per-pair set-up, writes the tile-ordered stream) and K3 (bulk-loads it and rasterises)
            rooflines.append(roofline_of("raster_tiles_kernel+tile_prims_kernel", raster_alg, k_ms["raster"],
                                         "vertex+index bytes + 80 B x instances + W x H x (1 B stencil + 16 B rgba32f)"))
        rooflines.sort(key=lambda r: -r["kernel_ms"])
        cfg = work.config_dict()
        cfg.update({"sharding": (("ONE render target tile-sharded over the ranks (16x16 tiles, owner (tx + ty) % N), finished tiles stored into every rank's attachments over NVLink"
                                  if args.config == 4 else
                                  "ONE render target composed from the ranks' contiguous draw-order slices: per tile the touching ranks hand the tile state from rank to rank over NVLink "
                                  "(same operations in the same order as one GPU), the last one stores the finished tile into every rank's attachments")
                                 if one_target else "independent scene per rank, no data-path collective"),
                    "frame_pipelining": (not args.no_pipelining),
                    "software_pipelining": ("the pass of step N is submitted after from_paths of step N + 1 (two batch objects alternate); each timed region ends with "
                                            "the last pass submitted and waited for") if software_pipelined else None,
                    "l2": "working set per step (target + vertex / index / record / pair arrays) exceeds the 126 MB L2 for configs 3-5; the target is cleared and re-written every step"})
        line = {
            "metric": "paths/sec", "value": total_paths * args.steps / (ms_dev * 1e-3), "unit": "paths/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "ms_per_step_one_at_a_time": ms_serial, "higher_is_better": True, "scaling": "strong" if one_target else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "covered_mpixel_per_s": total_covered * args.steps / (ms_dev * 1e-3) / 1e6,
            "covered_samples_per_step": total_covered,
            "e2e": {"value": total_paths * args.steps / (ms_e2e * 1e-3), "unit": "paths/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": d2h_result_bytes,
                    "ms_per_step": ms_e2e / args.steps,
                    "result_read": ("every step's pass counters are copied to the host inside the timed region; the host reads step N's while step N + 1 runs "
                                    "(frame pipelining) and waits for the last step's before the timed region ends") if pipelined_e2e else
                                   "the host waits for every step's pass counters before it starts the next step"},
            "gpu_launches": launches,
            "kernel_ms_per_step": k_ms,
            "kernel_ms_sum_over_step": (k_ms["tess"] + k_ms["bin"] + k_ms["raster"]) / (ms_dev / args.steps) if ms_dev > 0 else None,
            "kernel_ms_note": ("kernel_ms_per_step is measured with one step at a time (CUDA events inside the library); with frame pipelining the stages of "
                               "consecutive steps overlap, so their sum exceeds ms_per_step"),
            "roofline": rooflines[0],
            "roofline_other": rooflines[1:],
            "stages": {"tessellation": {"ms": k_ms["tess"], "algorithmic_bytes": tess_alg, "achieved_gbs": tess_alg / (k_ms["tess"] * 1e-3) / 1e9 if k_ms["tess"] else 0.0},
                       "binning": {"ms": k_ms["bin"], "tile_pairs": int(st.tile_pairs), "primitives": int(st.primitives)},
                       "raster": {"ms": k_ms["raster"]}},
            "clocks": clocks,
        }
        if frame_line is not None:
            frame_line["value"] = total_paths * args.steps / (frame_line["ms_per_step"] * args.steps * 1e-3)
            line["e2e_with_frame"] = frame_line
        if sharded is not None:
            line["tile_sharded_check"] = sharded
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_leg(args.config, args.scale)
        print(json.dumps(line), flush=True)
    if target is not None:
        target.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
