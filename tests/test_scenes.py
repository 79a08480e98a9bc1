"""Host logic of the scene generators for BASELINE configs 4 and 5 (CPU; the oracle is the checker)."""
import numpy as np

from contrast_renderer_b200 import scenes, sharding


def test_tiger_like_script_is_balanced_and_tessellates(oracle):
    """Config 4: every copy's script opens and closes its clips and opacity groups in nested order, stays inside the
    renderer's limits and ends at clip depth 0; instance slots are one per (copy, Shape); the group tessellates."""
    n = 5
    scene = scenes.tiger_like(n, extent=(640, 400))
    assert scene.paths.n_paths == 240 and scene.n_shapes == 24
    assert scene.transforms.shape == (n * 24, 16) and scene.colors.shape == (n * 24, 4)
    script = scene.script.astype(np.int64)
    per_copy = len(script) // n
    assert per_copy * n == len(script)
    for c in range(n):
        rows = script[c * per_copy:(c + 1) * per_copy]
        assert (rows[:, 1] == c * 24 + rows[:, 0]).all() and (rows[:, 2] == rows[:, 1] + 1).all()
        clip_stack, saved = [], []
        for shape, _, _, op, depth, save, restore in rows:
            if op == scenes.OP_CLIP:
                clip_stack.append(shape)
                assert depth == len(clip_stack) <= 3
            elif op == scenes.OP_UNCLIP:
                assert clip_stack.pop() == shape and depth == len(clip_stack)
            elif op == scenes.OP_SAVE_ALPHA:
                saved.append((shape, save))
                assert save == len(saved) - 1 < scene.alpha_layer_count
            elif op == scenes.OP_RESTORE_ALPHA:
                assert saved.pop() == (shape, restore)
            elif op in (scenes.OP_COLOR, scenes.OP_SCALE_ALPHA):
                assert depth == len(clip_stack)
            elif op == scenes.OP_STENCIL:   # the stencil draw that precedes an UnClip already uses the outer depth (src/renderer.rs:253-266)
                assert depth in (len(clip_stack), len(clip_stack) - 1)
        assert not clip_stack and not saved
    kinds = set(int(t) for t in scene.paths.segment_types)
    assert kinds == {0, 3, 4}, "lines, rational quadratics (constructors) and rational cubics (degree-elevated)"
    for i in range(scene.n_shapes):
        ref = oracle.shape_from_paths([], scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
        assert ref.vertex_offsets[7] > ref.vertex_offsets[6], "every Shape has a hull"


def test_config5_shapes_are_local(oracle):
    """Config 5: the 1 M-path scene is a grid of Shapes in LOCAL coordinates (scenes.py, MODEL UNITS): small coordinates,
    small hulls. With absolute 8K coordinates the reference's hull tolerance drowns in f32 noise (~1500 hull vertices)."""
    scene = scenes.dashed_rational_strokes(1000000)
    assert scene.paths.n_paths == 1000000 and scene.n_shapes == 1000 and scene.origins.shape == (1000, 2)
    assert float(np.abs(scene.paths.start).max()) < 16.0 and float(np.abs(scene.paths.segments[4][:, 4:]).max()) < 32.0
    px = scene.origins * scene.pixels_per_unit
    assert px[:, 0].min() > 0 and px[:, 0].max() < scene.width and px[:, 1].min() > 0 and px[:, 1].max() < scene.height
    part = sharding.shard_scene(scene, 8, 3)   # one rank's share of the 8-GPU configuration
    assert part.n_shapes == 125 and part.paths.n_paths == 125000
    for i in (0, 124):
        ref = oracle.shape_from_paths(part.dynamic_stroke_options, part.paths, int(part.shape_path_begin[i]), int(part.shape_path_begin[i + 1]))
        hull_vertices = (ref.vertex_offsets[7] - ref.vertex_offsets[6]) // 8
        assert 3 <= hull_vertices < 120, hull_vertices
        assert ref.proto_hull_points > 20000
