"""BASELINE config 5, one GPU's share at full target size: n dashed, round-joined, round-capped stroked paths of two
rational cubics each into a 7680x4320 target (the 8-GPU configuration gives each rank 125 000 of the 1 000 000 paths)."""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contrast_renderer_b200 import renderer as R, scenes, sharding
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8          # rank 0's share of the 1 M-path scene on `world` GPUs
t0 = time.time(); sc = scenes.dashed_rational_strokes(1000000)
if world > 1:
    sc = sharding.shard_scene(sc, world, 0)
n = sc.paths.n_paths
print(f"scene {n} paths in {time.time()-t0:.2f} s, {sc.n_shapes} shapes")
rnd = R.Renderer(); rnd.resize_internal_buffers(sc.width, sc.height); rnd.enable_timing(True)
cmds = scenes.stencil_cover_commands(sc.n_shapes)
batch = None
for it in range(3):
    t0 = time.time()
    batch = R.ShapeBatch(rnd, sc.dynamic_stroke_options, sc.paths, sc.shape_path_begin, existing=batch)
    rp = rnd.begin_render_pass(); rp.set_instances(sc.transforms(), sc.colors); rp.render_batch(batch, cmds); rp.submit()
    st = rnd.stats(); dt = time.time() - t0
    print(f"config5 {n} paths {sc.width}x{sc.height}: step {dt*1e3:.2f} ms (host clock) tess {st.last_tess_ms:.3f} (hull sort {st.last_hull_sort_ms:.3f} chain {st.last_hull_chain_ms:.3f}) bin {st.last_bin_ms:.3f} raster {st.last_raster_ms:.3f} ms; "
          f"vertex bytes {st.vertex_bytes/1e6:.1f} MB prims {st.primitives} pairs {st.tile_pairs} covered {st.covered_samples} proto {st.proto_hull_points} -> {n/dt/1e6:.2f} M paths/s")
