"""One small scene through the whole hot path, for compute-sanitizer runs: python tools/sanitize_probe.py fills|strokes|clip"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from contrast_renderer_b200 import renderer as R, scenes

kind = sys.argv[1] if len(sys.argv) > 1 else "fills"
scene = {"fills": lambda: scenes.mixed_fills(48, extent=(192, 128), rational=True),
         "strokes": lambda: scenes.closed_cubic_strokes(12, extent=(192, 128)),
         "dashes": lambda: scenes.dashed_rational_strokes(24, paths_per_shape=6, extent=(192, 128)),
         "text": lambda: scenes.glyph_like_fills(300, glyphs_per_shape=50, extent=(256, 128))}[kind]()
rnd = R.Renderer(R.Configuration(device=0))
rnd.resize_internal_buffers(scene.width, scene.height)
print("build", flush=True)
batch = R.ShapeBatch(rnd, scene.dynamic_stroke_options, scene.paths, scene.shape_path_begin)
print("built", flush=True)
rp = rnd.begin_render_pass()
rp.set_instances(scene.transforms(), scene.colors)
rp.render_batch(batch, scenes.stencil_cover_commands(scene.n_shapes))
rp.submit()
print("covered", int(rnd.stats().covered_samples), flush=True)
batch.close()
rnd.close()
