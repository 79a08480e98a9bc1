#!/usr/bin/env python
"""bench.py — paths/sec and covered-Mpixel/s of the tessellate -> stencil-then-cover hot path on the 100k-glyph 4K scene
(BASELINE.json configs[2]), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one whole pass of the hot path over the scene: Shape::from_paths for every shape of the scene (one batched
launch sequence) followed by one render pass (Stencil + Color per shape) into a cleared 3840x2160 target.
  value : paths/s with the path arrays already resident in HBM when the timed region starts
  e2e   : the same metric through the C-ABI with HOST (pinned) input arrays — host->device staging of every input inside
          the timed region — and a device->host read of the pass result (the covered-sample counter) every step
N > 1: path instances shard across GPUs by batch (north_star): each rank owns an independent scene of the same size
(weak scaling, no data-path collective); value = all paths of all ranks / max-over-ranks time.
`--impl reference`: the reference is a Rust crate with no toolchain in this image, so the reference arm is its CPU
restatement (oracle/, `kind: "port"`) on all host threads, on a bounded sample of the same scene.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = ("100k TTF glyph instances via the text front-end (paths_of_text layout, OpenSans outlines: 143.6k contour paths, 1.6M line + integral-quadratic "
            "segments), 12 px glyphs, 3840x2160, one Shape (Stencil+Color) per 160-glyph run")
N_GLYPHS = 100000
EXTENT = (3840, 2160)
RASTER_DRAM_BYTES_R01 = 218110976   # raster_tiles_kernel, one launch on this workload: 112.59 MB read + 105.52 MB written (ncu --set full, profiles/)
CHAIN_DRAM_BYTES_R01 = 21055744     # hull_chain_kernel, one launch: 21.06 MB read + 0 written
GLYPHS_PER_SHAPE = 160


def make_scene(rank: int, n_glyphs: int = N_GLYPHS):
    from contrast_renderer_b200 import scenes
    return scenes.text_glyphs(n_glyphs, seed=scenes.SEED0 + 3 + 1000 * rank, extent=EXTENT, glyphs_per_shape=GLYPHS_PER_SHAPE)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:   # driver-written: STREAM-style copy bandwidth of this pool's B200s (B200_PROFILING.md)
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except (OSError, KeyError, TypeError, ValueError):
        return 6650.0, "fallback"   # the recipe's stated fallback


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


def run_reference(args, rank: int, world: int):
    """The reference arm: the CPU restatement of the reference on all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle
    from contrast_renderer_b200 import scenes
    from contrast_renderer_b200.renderer import Configuration
    threads = oracle.max_threads()
    n_glyphs = args.ref_glyphs
    scene = make_scene(0, n_glyphs)
    cfg = Configuration().to_c()
    cmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in scenes.stencil_cover_commands(scene.n_shapes)]
    transforms = scene.transforms()

    def step():
        shapes = [oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
                  for i in range(scene.n_shapes)]
        _, _, _, covered = oracle.render(cfg, scene.width, scene.height, shapes, cmds, transforms, scene.colors, threads=threads)
        return covered

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    covered = 0
    for _ in range(args.steps):
        covered = step()
    dt = time.perf_counter() - t0
    value = scene.paths.n_paths * args.steps / dt
    sample = f"first {n_glyphs} glyphs ({scene.paths.n_paths} paths, {scene.n_shapes} shapes) of the workload, full 3840x2160 target; tessellation 1 thread (the reference loop is sequential, src/renderer.rs:187), raster {threads} threads"
    line = {
        "impl": "reference", "metric": "paths/sec", "value": value, "unit": "paths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "width": EXTENT[0], "height": EXTENT[1], "msaa_sample_count": 1, "sample_glyphs": n_glyphs},
        "covered_mpixel_per_s": covered * args.steps / dt / 1e6,
        "cpu_baseline": {"value": value, "unit": "paths/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(budget_s: float = 12.0):
    """Oracle (the CPU port of the reference) on a bounded sample of the same workload: sequential tessellation like the
    reference, raster on all host threads. Sample size is calibrated so the leg takes about `budget_s` seconds."""
    from oracle import oracle
    from contrast_renderer_b200 import scenes
    from contrast_renderer_b200.renderer import Configuration
    threads = oracle.max_threads()
    cfg = Configuration().to_c()

    def run(n_glyphs):
        scene = make_scene(0, n_glyphs)
        cmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in scenes.stencil_cover_commands(scene.n_shapes)]
        t0 = time.perf_counter()
        shapes = [oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
                  for i in range(scene.n_shapes)]
        t1 = time.perf_counter()
        oracle.render(cfg, scene.width, scene.height, shapes, cmds, scene.transforms(), scene.colors, threads=threads)
        t2 = time.perf_counter()
        return scene.paths.n_paths, t1 - t0, t2 - t1

    n = 2000
    paths, t_tess, t_raster = run(n)
    per_glyph = (t_tess + t_raster) / n
    n = int(min(N_GLYPHS, max(n, budget_s / max(per_glyph, 1e-9))))
    reps, tot_paths, tot_tess, tot_raster = 0, 0, 0.0, 0.0
    while reps < 12 and tot_tess + tot_raster < budget_s:   # the whole workload takes ~1.4 s on this host: repeat it to fill the budget
        paths, t_tess, t_raster = run(n)
        reps, tot_paths, tot_tess, tot_raster = reps + 1, tot_paths + paths, tot_tess + t_tess, tot_raster + t_raster
    return {"value": tot_paths / (tot_tess + tot_raster), "unit": "paths/s", "cores": threads, "kind": "port",
            "sample": f"{reps} x the first {n} glyphs ({paths} paths) of the workload into the full 3840x2160 target; tessellation on 1 thread "
                      f"{tot_tess / reps:.2f} s per pass (the reference loop is sequential, src/renderer.rs:187), raster on {threads} threads {tot_raster / reps:.2f} s per pass"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-glyphs", type=int, default=8000, help="sample size of one step of the reference arm")
    ap.add_argument("--glyphs", type=int, default=N_GLYPHS, help="debug only: a smaller scene is not the benchmark workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from contrast_renderer_b200 import _abi, renderer as R, scenes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    scene = make_scene(rank, args.glyphs)
    soa = scene.paths
    n_paths = soa.n_paths
    cmds = scenes.stencil_cover_commands(scene.n_shapes)
    transforms = scene.transforms()

    rnd = R.Renderer(R.Configuration(device=local_rank))
    # The renderer and the CUDA events that time it share ONE stream. It must not be the legacy default stream: its handle is
    # 0, which cr_renderer_set_stream reads as "use your own stream".
    stream = torch.cuda.Stream(dev)
    assert stream.cuda_stream != 0
    torch.cuda.set_stream(stream)
    rnd.set_stream(stream.cuda_stream)
    rnd.resize_internal_buffers(scene.width, scene.height)
    rnd.enable_timing(True)

    # device-resident and pinned-host copies of every input array
    host_arrays = soa.arrays()
    if not soa.any_stroked():
        host_arrays[9] = host_arrays[9][:0]   # all paths are filled: cr_path_soa.stroke_options = NULL, nothing to copy
    pinned = [torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).pin_memory() for a in host_arrays]
    resident = [t.to(dev) for t in pinned]
    inst_host = [torch.from_numpy(transforms.reshape(-1).copy()).pin_memory(), torch.from_numpy(scene.colors.reshape(-1).copy()).pin_memory()]
    inst_dev = [t.to(dev) for t in inst_host]
    h2d_bytes = sum(t.numel() for t in pinned) + sum(t.numel() * 4 for t in inst_host) + cmds.nbytes + 4 * len(scene.shape_path_begin)

    state = {"batch": None}

    def step(device_resident: bool):
        if device_resident:
            ptrs = [t.data_ptr() if t.numel() else 0 for t in resident]
            space = _abi.CR_MEM_DEVICE
        else:
            ptrs = [t.data_ptr() if t.numel() else 0 for t in pinned]
            space = _abi.CR_MEM_HOST
        state["batch"] = R.ShapeBatch(rnd, scene.dynamic_stroke_options, soa, scene.shape_path_begin, existing=state["batch"], memory_space=space, pointers=ptrs)
        rp = rnd.begin_render_pass()
        if device_resident:
            rp.set_instances(inst_dev[0].data_ptr(), inst_dev[1].data_ptr(), count=scene.n_shapes, memory_space=_abi.CR_MEM_DEVICE)
        else:
            rp.set_instances(inst_host[0].data_ptr(), inst_host[1].data_ptr(), count=scene.n_shapes, memory_space=_abi.CR_MEM_HOST)
        rp.render_batch(state["batch"], cmds)
        rp.submit()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(device_resident: bool, steps: int, collect: bool):
        kernel_ms = {"tess": [], "bin": [], "raster": [], "hull_sort": [], "hull_chain": []}
        covered = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step(device_resident)
            if collect or not device_resident:
                st = rnd.stats()   # device->host read of the pass result (covered-sample counter); synchronises the stream
                covered = int(st.covered_samples)
                kernel_ms["tess"].append(st.last_tess_ms)
                kernel_ms["bin"].append(st.last_bin_ms)
                kernel_ms["raster"].append(st.last_raster_ms)
                kernel_ms["hull_sort"].append(st.last_hull_sort_ms)
                kernel_ms["hull_chain"].append(st.last_hull_chain_ms)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, covered, kernel_ms

    # warm-up (both arms), then the two timed regions
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    import time
    t_warm = time.time()
    for _ in range(max(3, args.warmup)):
        step(True)
    step(False)
    while rank == 0 and time.time() - t_warm < 0.5:   # untimed: gives nvidia-smi time to come up, under the benchmark's own load
        step(True)
    torch.cuda.synchronize(dev)
    launches0 = int(rnd.stats().kernel_launches)
    sampler.mark_begin()
    ms_dev, covered, kernel_ms = timed(True, args.steps, collect=True)
    launches = int(rnd.stats().kernel_launches) - launches0
    ms_e2e, covered_e2e, _ = timed(False, args.steps, collect=False)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    st = rnd.stats()

    total_paths, total_covered = n_paths, covered
    if world > 1:
        t = torch.tensor([n_paths, covered], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_paths, total_covered = int(t[0].item()), int(t[1].item())

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        layout_bytes = int(st.vertex_bytes)
        fb_bytes = scene.width * scene.height * (1 + 16)
        raster_alg = layout_bytes + 80 * scene.n_shapes + fb_bytes           # SURVEY §8d B_rast
        tess_alg = int(st.input_bytes) + layout_bytes                         # SURVEY §8d B_tess
        mean = lambda xs: float(np.mean(xs)) if xs else 0.0
        k_ms = {k: mean(v) for k, v in kernel_ms.items()}
        # Roofline of the DOMINANT single kernel, whichever of the two heavy ones took longer in this run (both are timed live
        # with CUDA events on the renderer's stream; the committed ncu launch list profiles/launches_r01.csv shows the same
        # shares): K3 raster_tiles_kernel with B_rast of SURVEY 8d, or hull_chain_kernel (convex_hull::andrew's chains) with
        # 8 B per proto-hull point read + 8 B per hull vertex written. The other one is reported
        # next to it as "roofline_other".
        proto_points, hull_vertices = int(st.proto_hull_points), int(st.hull_vertices)
        chain_alg = 8 * proto_points + 8 * hull_vertices

        def roofline_of(name, alg, ms, formula, traffic):
            ach = alg / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak if peak else None,
                    "algorithmic_bytes": alg, "algorithmic_bytes_formula": formula, "kernel_ms": ms, "traffic": traffic}
        default_workload = args.glyphs == N_GLYPHS
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/ncu_full_r01_summary.csv)
        rooflines = [roofline_of("raster_tiles_kernel", raster_alg, k_ms["raster"], "vertex+index bytes + 80 B x instances + W x H x (1 B stencil + 16 B rgba32f)",
                                 RASTER_DRAM_BYTES_R01 if default_workload else None),
                     roofline_of("hull_chain_kernel", chain_alg, k_ms["hull_chain"], "8 B x proto-hull points + 8 B x hull vertices (latency-bound sequential stack machines, DESIGN.md section 7)",
                                 CHAIN_DRAM_BYTES_R01 if default_workload else None)]
        rooflines.sort(key=lambda r: -r["kernel_ms"])
        line = {
            "metric": "paths/sec", "value": total_paths * args.steps / (ms_dev * 1e-3), "unit": "paths/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "paths_per_gpu": n_paths, "segments_per_gpu": soa.n_segments, "shapes_per_gpu": scene.n_shapes,
                       "width": scene.width, "height": scene.height, "msaa_sample_count": 1, "winding_counter_bits": 4, "clip_nesting_counter_bits": 4,
                       "color_format": "rgba32f", "sharding": "independent scene per rank, no collective",
                       "l2": "working set per step (141 MB target + vertex/index/pair arrays) exceeds the 126 MB L2; the target is cleared and re-written every step"},
            "covered_mpixel_per_s": total_covered * args.steps / (ms_dev * 1e-3) / 1e6,
            "covered_samples_per_step": total_covered,
            "e2e": {"value": total_paths * args.steps / (ms_e2e * 1e-3), "unit": "paths/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 8 + 4 * 12 + 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "kernel_ms_per_step": k_ms,
            "roofline": rooflines[0],
            "roofline_other": rooflines[1],
            "stages": {"tessellation": {"ms": k_ms["tess"], "algorithmic_bytes": tess_alg, "achieved_gbs": tess_alg / (k_ms["tess"] * 1e-3) / 1e9 if k_ms["tess"] else 0.0},
                       "binning": {"ms": k_ms["bin"], "tile_pairs": int(st.tile_pairs), "primitives": int(st.primitives)},
                       "raster": {"ms": k_ms["raster"]}},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
