"""Regression fixture: SHA-256 of the oracle's outputs (vertex / index / descriptor buffers, stencil and colour frames) for a
few seeded scenes. These are NOT reference vectors (the reference cannot run here); they freeze the
oracle — the checker every parity test relies on — so that an edit that changes its behaviour cannot go unnoticed.
"inputs" is the hash of the generated scene itself (numpy's vectorised sin / cos may differ in the last bit between CPU
generations; the test skips a scene whose inputs differ instead of blaming the oracle). Regenerate deliberately with:  python tests/golden/make_oracle_hashes.py --write"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from contrast_renderer_b200 import scenes  # noqa: E402
from contrast_renderer_b200.renderer import Configuration  # noqa: E402

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_hashes.json")


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def compute(oracle) -> dict:
    out = {}
    cases = {
        "mixed_fills_all_kinds": (scenes.mixed_fills(40, extent=(200, 150), rational=True, seed=5), Configuration()),
        "closed_cubic_strokes": (scenes.closed_cubic_strokes(16, extent=(200, 150)), Configuration()),
        "dashed_rational_strokes_msaa4": (scenes.dashed_rational_strokes(60, extent=(200, 150), paths_per_shape=10, pixels_per_unit=4.0),
                                          Configuration(msaa_sample_count=4)),
    }
    for name, (scene, config) in cases.items():
        shapes = [oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
                  for i in range(scene.n_shapes)]
        cmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in scenes.stencil_cover_commands(scene.n_shapes)]
        color, stencil, _, covered = oracle.render(config.to_c(), scene.width, scene.height, shapes, cmds, scene.transforms(), scene.colors)
        out[name] = {"inputs": digest(*scene.paths.arrays(), scene.transforms(), scene.colors),
                     "tessellation": digest(*[s.vertex_buffer for s in shapes], *[s.index_buffer for s in shapes], *[s.stroke_buffer for s in shapes]),
                     "frame": digest(color, stencil), "covered_samples": int(covered)}
    tiger = scenes.tiger_like(3, extent=(200, 150), instance_px=(60.0, 140.0))
    shapes = [oracle.shape_from_paths([], tiger.paths, int(tiger.shape_path_begin[i]), int(tiger.shape_path_begin[i + 1])) for i in range(tiger.n_shapes)]
    color, stencil, layers, covered = oracle.render(Configuration(alpha_layer_count=2).to_c(), tiger.width, tiger.height, shapes, tiger.oracle_commands(),
                                                    tiger.transforms, tiger.colors)
    out["tiger_like_clips_and_opacity"] = {"inputs": digest(*tiger.paths.arrays(), tiger.transforms, tiger.colors, tiger.script),
                                           "tessellation": digest(*[s.vertex_buffer for s in shapes], *[s.index_buffer for s in shapes]),
                                           "frame": digest(color, stencil, *layers), "covered_samples": int(covered)}
    return out


if __name__ == "__main__":
    from oracle import oracle
    oracle.build()
    result = compute(oracle)
    if "--write" in sys.argv:
        with open(PATH, "w") as f:
            json.dump(result, f, indent=1, sort_keys=True)
    print(json.dumps(result, indent=1, sort_keys=True))
