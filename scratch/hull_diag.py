import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contrast_renderer_b200 import renderer as R, scenes
sc = scenes.text_glyphs()
rnd = R.Renderer(); rnd.resize_internal_buffers(sc.width, sc.height)
batch = R.ShapeBatch(rnd, sc.dynamic_stroke_options, sc.paths, sc.shape_path_begin)
rnd.synchronize()
print("done", sc.n_shapes)
