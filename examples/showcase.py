#!/usr/bin/env python
"""The reference's demo (examples/showcase/main.rs) on a B200, without a window: one frame to a PNG.

    python examples/showcase.py [out.png] [--size 1280x720] [--turn 0.6,0.25] [--distance 5]

Builds the demo's Shape ("Hello World" + a dashed rounded rectangle) with the host mirrors of the crate's helpers
(`text.paths_of_text`, `Path::from_rounded_rect`, `Path::reverse`, the ppga3d camera motors of `utils.rs`), tessellates and renders
it through the C-ABI exactly like the demo does (4x MSAA, depth LessEqual + write, back-face culling, Stencil + Color per
instance), resolves the four samples and writes the frame."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contrast_renderer_b200 import renderer as R, scenes, utils  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out", nargs="?", default="showcase.png")
    ap.add_argument("--size", default="1280x720")
    ap.add_argument("--turn", default="0.0,0.0", help="camera rotation about the y and x axes in radians (the demo: mouse position)")
    ap.add_argument("--distance", type=float, default=5.0, help="camera distance (the demo: mouse wheel, 2..100)")
    args = ap.parse_args()
    width, height = (int(v) for v in args.size.split("x"))
    turn = tuple(float(v) for v in args.turn.split(","))
    paths, shape_path_begin, dynamic_stroke_options, transforms, colors = scenes.showcase((width, height), turn, args.distance)
    rnd = R.Renderer(R.Configuration(msaa_sample_count=4, depth_compare=R.CompareFunction.LessEqual, depth_write_enabled=True,
                                     cull_mode=R.CullMode.Back))
    rnd.resize_internal_buffers(width, height)
    shape = R.Shape.from_paths(rnd, dynamic_stroke_options, paths)
    rp = rnd.begin_render_pass()
    rp.set_instances(transforms, colors)
    for i in range(len(transforms)):
        shape.render(rp, range(i, i + 1), R.RenderOperation.Stencil)
        shape.render(rp, range(i, i + 1), R.RenderOperation.Color)
    rp.submit()
    frame = rnd.read_color().mean(axis=2)   # the demo's resolve target
    stats = rnd.stats()
    utils.save_png(args.out, frame)
    print(f"{args.out}: {width}x{height}, {int(stats.primitives)} candidate primitives, {int(stats.tile_pairs)} (tile, primitive) pairs, "
          f"{int(stats.covered_samples)} covered samples")
    shape.close()
    rnd.close()


if __name__ == "__main__":
    main()
