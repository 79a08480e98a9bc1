"""BASELINE.json configs 2, 4 and 5 at (or near) their full sizes: the CUDA path through the C-ABI against the CPU oracle,
whole frames bit for bit (u8 stencil, f32 colour bit patterns, alpha layers, covered-sample counts) plus sampled Shapes'
vertex / index buffers. Configs 1 and 3 at full size live in test_parity_gpu.py. Oracle time on the 16-thread GPU box:
config 2 ~2 s per fill rule, config 4 (1000 copies, 4K) ~6 s, config 5 (100 k of the 1 M paths, 8K) ~8 s."""
import hashlib

import numpy as np
import pytest

from contrast_renderer_b200 import scenes

pytestmark = pytest.mark.gpu


def oracle_shapes(oracle, scene):
    return [oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
            for i in range(scene.n_shapes)]


def digest(array: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(array).view(np.uint8).tobytes()).hexdigest()


def assert_shapes_equal(batch, refs, which):
    for i in which:
        layout = batch[i].layout()
        assert list(layout.vertex_offsets) == refs[i].vertex_offsets and list(layout.index_offsets) == refs[i].index_offsets, f"shape {i}: layout"
        assert np.array_equal(batch[i].vertex_buffer(), refs[i].vertex_buffer), f"shape {i}: vertex bytes"
        assert np.array_equal(batch[i].index_buffer(), refs[i].index_buffer), f"shape {i}: index bytes"


def assert_frame_equal(color, stencil, ref_color, ref_stencil):
    # frame hashes first (cheap); the element-wise comparison only runs to say WHERE a mismatch is
    if digest(stencil) != digest(ref_stencil):
        raise AssertionError(f"stencil differs at {np.argwhere(stencil != ref_stencil)[:5]}")
    if digest(color) != digest(ref_color):
        raise AssertionError(f"colour differs at {np.argwhere(color.view(np.uint32) != ref_color.view(np.uint32))[:5]}")


@pytest.mark.parametrize("winding_bits", [1, 4], ids=["even_odd", "non_zero"])
def test_config2_full_size(cr, oracle, winding_bits):
    """10 000 mixed line / quadratic / cubic filled paths, even-odd and non-zero winding, 1920x1080."""
    scene = scenes.mixed_fills(10000)
    assert scene.paths.n_paths == 10000 and (scene.width, scene.height) == (1920, 1080)
    cfg = cr.Configuration(winding_counter_bits=winding_bits, clip_nesting_counter_bits=4)
    rnd = cr.Renderer(cfg)
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, scene.dynamic_stroke_options, scene.paths, scene.shape_path_begin)
    cmds = scenes.stencil_cover_commands(scene.n_shapes)
    rp = rnd.begin_render_pass()
    rp.set_instances(scene.transforms(), scene.colors)
    rp.render_batch(batch, cmds)
    rp.submit()
    color, stencil, covered = rnd.read_color(), rnd.read_stencil(), int(rnd.stats().covered_samples)
    refs = oracle_shapes(oracle, scene)
    assert_shapes_equal(batch, refs, range(0, scene.n_shapes, 97))
    ocmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in cmds]
    ref_color, ref_stencil, _, ref_covered = oracle.render(cfg.to_c(), scene.width, scene.height, refs, ocmds, scene.transforms(), scene.colors,
                                                           threads=oracle.max_threads())
    assert ref_covered > 50_000_000
    assert_frame_equal(color, stencil, ref_color, ref_stencil)
    assert covered == ref_covered
    batch.close()
    rnd.close()


def test_config4_full_size(cr, oracle):
    """1000 placed copies of the 240-path constructor-built group, three nested clips and two nested opacity groups per copy,
    3840x2160 (single GPU here; the tile-sharded 4-GPU run of the same frame is tests/multi_gpu/tile_sharding_check.py)."""
    scene = scenes.tiger_like(1000)
    assert (scene.width, scene.height) == (3840, 2160) and scene.paths.n_paths == 240 and len(scene.script) == 60000
    config = cr.Configuration(alpha_layer_count=2)
    rnd = cr.Renderer(config)
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
    refs = oracle_shapes(oracle, scene)
    assert_shapes_equal(batch, refs, range(scene.n_shapes))
    rp = rnd.begin_render_pass()
    rp.set_instances(scene.transforms, scene.colors)
    scene.record(rp, batch)
    rp.submit()
    color, stencil, covered = rnd.read_color(), rnd.read_stencil(), int(rnd.stats().covered_samples)
    layers = [rnd.read_alpha_layer(k) for k in range(2)]
    ref_color, ref_stencil, ref_layers, ref_covered = oracle.render(config.to_c(), scene.width, scene.height, refs, scene.oracle_commands(),
                                                                    scene.transforms, scene.colors, threads=oracle.max_threads())
    assert ref_covered > 30_000_000
    assert_frame_equal(color, stencil, ref_color, ref_stencil)
    for k in range(2):
        assert digest(layers[k]) == digest(ref_layers[k]), f"alpha layer {k}"
    assert covered == ref_covered
    batch.close()
    rnd.close()


def test_config5_100k_paths_8k(cr, oracle):
    """100 000 of config 5's 1 000 000 dashed, round-joined, round-capped strokes of two rational cubics each, into the full
    7680x4320 target (one rank's share of the 8-GPU configuration is 125 000 paths)."""
    scene = scenes.dashed_rational_strokes(100000)
    assert (scene.width, scene.height) == (7680, 4320) and scene.paths.n_paths == 100000
    rnd = cr.Renderer()
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, scene.dynamic_stroke_options, scene.paths, scene.shape_path_begin)
    cmds = scenes.stencil_cover_commands(scene.n_shapes)
    rp = rnd.begin_render_pass()
    rp.set_instances(scene.transforms(), scene.colors)
    rp.render_batch(batch, cmds)
    rp.submit()
    color, stencil, covered = rnd.read_color(), rnd.read_stencil(), int(rnd.stats().covered_samples)
    refs = oracle_shapes(oracle, scene)
    assert_shapes_equal(batch, refs, range(0, scene.n_shapes, 9))
    ocmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in cmds]
    ref_color, ref_stencil, _, ref_covered = oracle.render(rnd.config.to_c(), scene.width, scene.height, refs, ocmds, scene.transforms(), scene.colors,
                                                           threads=oracle.max_threads())
    assert ref_covered > 10_000_000
    assert_frame_equal(color, stencil, ref_color, ref_stencil)
    assert covered == ref_covered
    batch.close()
    rnd.close()
