import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contrast_renderer_b200 import renderer as R, scenes
sc = scenes.glyph_like_fills(100000)
rnd = R.Renderer(); rnd.resize_internal_buffers(sc.width, sc.height)
cmds = scenes.stencil_cover_commands(sc.n_shapes)
batch = None
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    batch = R.ShapeBatch(rnd, sc.dynamic_stroke_options, sc.paths, sc.shape_path_begin, existing=batch)
    rp = rnd.begin_render_pass(); rp.set_instances(sc.transforms(), sc.colors); rp.render_batch(batch, cmds); rp.submit()
    rnd.synchronize()
print("covered", rnd.stats().covered_samples)
