"""Parity tests proper: the CUDA path, called through the C-ABI, against the CPU oracle on the same seeded inputs.
Bit-exact for everything (vertex bytes, u16 indices, 48-byte descriptors, u8 stencil, f32 colour bit patterns): both
sides compute in f32 with the shared arithmetic contract (csrc/arith/cr_arith.h), FMA contraction off."""
import numpy as np
import pytest

from contrast_renderer_b200 import _abi, scenes
from contrast_renderer_b200.path import (Cap, CurveApproximation, DashInterval, DynamicStrokeOptions, Join, Path, PathSoA,
                                         RationalCubicCurveSegment, RationalQuadraticCurveSegment, StrokeOptions)

pytestmark = pytest.mark.gpu


def oracle_shapes(oracle, scene):
    return [oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
            for i in range(scene.n_shapes)]


def assert_shape_equal(oracle, shape, ref, tag=""):
    layout = shape.layout()
    assert list(layout.vertex_offsets) == ref.vertex_offsets, f"{tag}: vertex_offsets"
    assert list(layout.index_offsets) == ref.index_offsets, f"{tag}: index_offsets"
    assert int(layout.proto_hull_points) == ref.proto_hull_points, f"{tag}: proto hull size"
    got_v, want_v = shape.vertex_buffer(), ref.vertex_buffer
    if not np.array_equal(got_v, want_v):
        got, want = oracle.split_vertex_buffer(got_v, ref.vertex_offsets), oracle.split_vertex_buffer(want_v, ref.vertex_offsets)
        for name, g, w in zip(oracle.CATEGORY_NAMES, got, want):
            if g.tobytes() != w.tobytes():
                bad = [i for i in range(len(g)) if g[i].tobytes() != w[i].tobytes()]
                raise AssertionError(f"{tag}: {name} vertices differ at {bad[:5]} of {len(g)}: got {g[bad[0]]} want {w[bad[0]]}")
    assert np.array_equal(shape.index_buffer(), ref.index_buffer), f"{tag}: index buffer"
    assert np.array_equal(shape.stroke_buffer(), ref.stroke_buffer), f"{tag}: stroke descriptors"


def render_both(cr, oracle, scene, config=None, commands=None, transforms=None, colors=None, per_command_state=None):
    """Tessellates + renders `scene` on the GPU and with the oracle; returns ((color, stencil, covered), (ref...))."""
    config = config or cr.Configuration()
    rnd = cr.Renderer(config)
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, scene.dynamic_stroke_options, scene.paths, scene.shape_path_begin)
    refs = oracle_shapes(oracle, scene)
    cmds = scenes.stencil_cover_commands(scene.n_shapes) if commands is None else np.asarray(commands, np.uint32)
    transforms = scene.transforms() if transforms is None else transforms
    colors = scene.colors if colors is None else colors
    rp = rnd.begin_render_pass()
    rp.set_instances(transforms, colors)
    rp.render_batch(batch, cmds)
    rp.submit()
    got = (rnd.read_color(), rnd.read_stencil(), int(rnd.stats().covered_samples))
    ocmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in cmds]
    ref_color, ref_stencil, _, ref_covered = oracle.render(config.to_c(), scene.width, scene.height, refs, ocmds, transforms, colors, threads=4)
    batch.close()
    rnd.close()
    return got, (ref_color, ref_stencil, ref_covered)


def assert_frames_equal(got, want):
    (color, stencil, covered), (ref_color, ref_stencil, ref_covered) = got, want
    assert np.array_equal(stencil, ref_stencil), f"stencil differs at {np.argwhere(stencil != ref_stencil)[:5]}"
    diff = color.view(np.uint32) != ref_color.view(np.uint32)
    assert not diff.any(), f"colour differs at {np.argwhere(diff)[:5]}"
    assert covered == ref_covered


# ----------------------------------------------------------------------------------------------- tessellation
@pytest.mark.parametrize("maker", [
    lambda: scenes.closed_cubic_strokes(200),                                   # BASELINE config 1 (reduced count)
    lambda: scenes.mixed_fills(600),                                            # config 2 kinds: line / quad / cubic
    lambda: scenes.mixed_fills(600, rational=True, seed=77),                    # all five segment kinds
    lambda: scenes.mixed_fills(300, rational=True, paths_per_shape=7, seed=5),  # multi-path shapes (hull over many paths)
    lambda: scenes.glyph_like_fills(1500, glyphs_per_shape=100),                # config 3 kinds
    lambda: scenes.dashed_rational_strokes(300, paths_per_shape=25),            # config 5 kinds
], ids=["cubic_strokes", "mixed_fills", "all_kinds", "multi_path_shapes", "glyphs", "dashed_rational_strokes"])
def test_tessellation_matches_oracle(cr, oracle, maker):
    scene = maker()
    rnd = cr.Renderer()
    batch = cr.ShapeBatch(rnd, scene.dynamic_stroke_options, scene.paths, scene.shape_path_begin)
    assert len(batch) == scene.n_shapes
    for i, ref in enumerate(oracle_shapes(oracle, scene)):
        assert_shape_equal(oracle, batch[i], ref, f"{scene.name} shape {i}")
    batch.close()
    rnd.close()


def test_config1_full_size_tessellation(cr, oracle):
    """BASELINE config 1 at its full size: 1k closed cubic paths, stroke tessellation to vertex buffers only."""
    scene = scenes.closed_cubic_strokes(1000)
    rnd = cr.Renderer()
    batch = cr.ShapeBatch(rnd, scene.dynamic_stroke_options, scene.paths, scene.shape_path_begin)
    for i, ref in enumerate(oracle_shapes(oracle, scene)):
        assert_shape_equal(oracle, batch[i], ref, f"shape {i}")
    batch.close()
    rnd.close()


def stroke_path(points_and_kinds, start, **so):
    p = Path(start, StrokeOptions(**so))
    for kind, data in points_and_kinds:
        if kind == "L":
            p.push_line(data)
        elif kind == "Q":
            p.push_integral_quadratic_curve(data)
        elif kind == "C":
            p.push_integral_cubic_curve(data)
        elif kind == "RQ":
            p.push_rational_quadratic_curve(RationalQuadraticCurveSegment(*data))
        else:
            p.push_rational_cubic_curve(RationalCubicCurveSegment(*data))
    return p


STROKE_VARIANTS = [
    dict(width=0.3, offset=0.0, miter_clip=1.0, closed=False),
    dict(width=0.3, offset=0.5, miter_clip=4.0, closed=True),
    dict(width=0.2, offset=-0.25, miter_clip=0.6, closed=True, curve_approximation=CurveApproximation.UniformlySpacedParameters(7)),
    dict(width=0.5, offset=0.1, miter_clip=2.0, closed=False, curve_approximation=CurveApproximation.UniformTangentAngle(0.35)),
]


@pytest.mark.parametrize("so", STROKE_VARIANTS, ids=["open", "closed_offset", "uniform_params", "coarse_angle"])
def test_single_shape_api_strokes_every_segment_kind(cr, oracle, so):
    """cr_shape_from_paths (Shape::from_paths) on hand-built paths: every segment kind, straight runs (no join),
    reversals (anti-parallel join), cusps of an S-cubic, both curve approximations, offsets, open and closed."""
    w = 0.70710678
    paths = [
        stroke_path([("L", [2, 0]), ("L", [2, 1]), ("L", [0, 1])], [0, 0], **so),                          # rectangle corners
        stroke_path([("L", [1, 0]), ("L", [2, 0]), ("L", [1, 0])], [0, 0], **so),                          # collinear then reversal
        stroke_path([("Q", [[1, 1.5], [2, 0]]), ("C", [[3, -1], [4, 2], [5, 0]])], [0, 0], **so),          # quad + S cubic (inflection)
        stroke_path([("RQ", (w, [[1, 1], [0, 1]])), ("RQ", (w, [[-1, 1], [-1, 0]]))], [1, 0], **so),       # two quarter circles
        stroke_path([("RC", ([1, 0.6, 1.7, 1], [[0.5, 1.5], [2.5, 1.2], [3, 0]])), ("L", [3, -1])], [0, 0], **so),
        stroke_path([("C", [[3, 2], [-1, 2], [2, 0]])], [0, 0], **so),                                     # loop cubic
    ]
    dso = [DynamicStrokeOptions.Solid(Join.Round, Cap.Round, Cap.Square)]
    soa = PathSoA.from_paths(paths)
    rnd = cr.Renderer()
    shape = cr.Shape.from_paths(rnd, dso, soa)
    assert_shape_equal(oracle, shape, oracle.shape_from_paths(dso, soa), "strokes")
    shape.close()
    rnd.close()


def test_mixed_stroke_and_fill_in_one_shape(cr, oracle):
    fill = Path([0, 0])
    fill.push_line([3, 0])
    fill.push_integral_quadratic_curve([[3.5, 1.5], [2, 2]])
    fill.push_integral_cubic_curve([[1.5, 3], [0.5, 1], [0, 2]])
    fill.close()
    stroke = stroke_path([("L", [1, 2]), ("Q", [[2, 3], [3, 1]])], [0.5, 0.5], width=0.2, closed=True, dynamic_stroke_options_group=1)
    dso = [DynamicStrokeOptions.Solid(Join.Miter, Cap.Butt, Cap.Butt),
           DynamicStrokeOptions.Dashed(Join.Bevel, [DashInterval(0.5, 1.0, Cap.Out, Cap.In), DashInterval(1.5, 2.0, Cap.Left, Cap.Right)], 0.25)]
    soa = PathSoA.from_paths([fill, stroke, fill])
    rnd = cr.Renderer()
    shape = cr.Shape.from_paths(rnd, dso, soa)
    ref = oracle.shape_from_paths(dso, soa)
    assert_shape_equal(oracle, shape, ref, "mixed")
    # Shape::set_dynamic_stroke_options: in-place 48-byte update, no re-tessellation
    new = DynamicStrokeOptions.Dashed(Join.Round, [DashInterval(0.1, 0.2, Cap.Round, Cap.Round)], 0.75)
    shape.set_dynamic_stroke_options(1, new)
    ref.set_dynamic_stroke_options(1, new)
    assert np.array_equal(shape.stroke_buffer(), ref.stroke_buffer)
    with pytest.raises(cr.DynamicStrokeOptionsIndexOutOfBounds):
        shape.set_dynamic_stroke_options(2, new)
    # existing shape is consumed and rebuilt in place
    soa2 = PathSoA.from_paths([stroke, fill])
    shape2 = cr.Shape.from_paths(rnd, dso, soa2, existing=shape)
    assert_shape_equal(oracle, shape2, oracle.shape_from_paths(dso, soa2), "rebuilt")
    shape2.close()
    rnd.close()


def test_edge_cases(cr, oracle):
    """Empty inputs, paths without segments, a single point, fewer than three hull points."""
    rnd = cr.Renderer()
    empty = PathSoA.from_paths([])
    shape = cr.Shape.from_paths(rnd, [], empty)
    assert_shape_equal(oracle, shape, oracle.shape_from_paths([], empty), "no paths")
    shape.close()
    lonely = PathSoA.from_paths([Path([1, 2])])                       # a filled path with no segments
    shape = cr.Shape.from_paths(rnd, [], lonely)
    assert_shape_equal(oracle, shape, oracle.shape_from_paths([], lonely), "no segments")
    shape.close()
    two = Path([0, 0])
    two.push_line([1, 1])
    soa = PathSoA.from_paths([two])
    shape = cr.Shape.from_paths(rnd, [], soa)
    assert_shape_equal(oracle, shape, oracle.shape_from_paths([], soa), "two points")
    shape.close()
    # ragged batch: empty shapes between populated ones
    scene = scenes.mixed_fills(40, rational=True)
    begin = np.array([0, 0, 13, 13, 13, 40, 40], np.uint32)
    batch = cr.ShapeBatch(rnd, [], scene.paths, begin)
    for i in range(len(begin) - 1):
        assert_shape_equal(oracle, batch[i], oracle.shape_from_paths([], scene.paths, int(begin[i]), int(begin[i + 1])), f"ragged {i}")
    batch.close()
    rnd.close()


def test_hull_paths_beyond_the_shared_memory_fast_path(cr, oracle):
    """convex_hull::andrew on the device has three regimes: shared-memory sort + chains, chains deeper than the
    shared-memory stacks (a hull with thousands of vertices), and shapes too large for shared memory (global sort)."""
    rnd = cr.Renderer()
    # (1) a 4000-gon of radius 200 (turn areas around ERROR_MARGIN): ~2600 proto-hull points stay hull vertices -> the
    # 1024-entry shared-memory chain stacks overflow and the chains are redone with global stacks
    ang = np.linspace(0.0, 2.0 * np.pi, 4000, endpoint=False)
    ring = Path([200.0, 0.0])
    for a in ang[1:]:
        ring.push_line([200.0 * np.cos(a), 200.0 * np.sin(a)])
    ring.close()
    soa = PathSoA.from_paths([ring])
    shape = cr.Shape.from_paths(rnd, [], soa)
    ref = oracle.shape_from_paths([], soa)
    assert ref.vertex_offsets[7] - ref.vertex_offsets[6] > 8 * 2200   # at least one chain is deeper than 1024
    assert_shape_equal(oracle, shape, ref, "4000-gon")
    shape.close()
    # (1b) a 1200-gon of radius 100: both chains (~600 entries) stay in the shared-memory stacks
    ang = np.linspace(0.0, 2.0 * np.pi, 1200, endpoint=False)
    ring = Path([100.0, 0.0])
    for a in ang[1:]:
        ring.push_line([100.0 * np.cos(a), 100.0 * np.sin(a)])
    ring.close()
    soa = PathSoA.from_paths([ring])
    shape = cr.Shape.from_paths(rnd, [], soa)
    ref = oracle.shape_from_paths([], soa)
    assert_shape_equal(oracle, shape, ref, "1200-gon")
    shape.close()
    # (2) one shape with ~48k proto-hull points (2000 glyphs): sorted in global memory
    scene = scenes.glyph_like_fills(4000, extent=(3840, 512), glyphs_per_shape=2000)
    batch = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
    for i in range(scene.n_shapes):
        ref = oracle.shape_from_paths([], scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
        assert ref.proto_hull_points > 26624
        assert_shape_equal(oracle, batch[i], ref, f"big shape {i}")
    batch.close()
    rnd.close()


def test_error_codes(cr, oracle):
    """Error behaviour of the reference API (src/error.rs:5-16, src/renderer.rs:32,189,366,433,933,947)."""
    with pytest.raises(cr.NumberOfStencilBitsIsUnsupported):
        cr.Renderer(cr.Configuration(winding_counter_bits=0))
    with pytest.raises(cr.NumberOfStencilBitsIsUnsupported):
        cr.Renderer(cr.Configuration(winding_counter_bits=5, clip_nesting_counter_bits=4))
    rnd = cr.Renderer(cr.Configuration(clip_nesting_counter_bits=2, winding_counter_bits=6, alpha_layer_count=1))
    p = stroke_path([("L", [1, 0])], [0, 0], width=0.1, dynamic_stroke_options_group=1)
    soa = PathSoA.from_paths([p])
    with pytest.raises(cr.DynamicStrokeOptionsIndexOutOfBounds):
        cr.Shape.from_paths(rnd, [DynamicStrokeOptions.Solid(Join.Miter, Cap.Butt, Cap.Butt)], soa)
    five = DynamicStrokeOptions.Dashed(Join.Miter, [DashInterval(i, i + 0.5) for i in range(5)], 0.0)
    with pytest.raises(cr.TooManyDashIntervals):
        cr.Shape.from_paths(rnd, [five, five], soa)
    with pytest.raises(cr.Error) as e:
        rnd.begin_render_pass()
    assert e.value.status == _abi.CR_ERR_NOT_RESIZED
    rnd.resize_internal_buffers(64, 64)
    rp = rnd.begin_render_pass()
    rp.set_clip_depth(3)
    with pytest.raises(cr.ClipStackOverflow):
        rp.set_clip_depth(4)
    rp.save_alpha_context(0)
    with pytest.raises(cr.TooManyNestedOpacityGroups):
        rp.save_alpha_context(1)
    with pytest.raises(cr.TooManyNestedOpacityGroups):
        rp.restore_alpha_context(1)
    rp.submit()
    # a curve with more tangent-angle steps than the device admits (2^22 per interval) is reported, not truncated silently
    tight = stroke_path([("Q", [[1, 2], [2, 0]])], [0, 0], width=0.1, curve_approximation=CurveApproximation.UniformTangentAngle(1e-7))
    with pytest.raises(cr.Error) as e:
        cr.Shape.from_paths(rnd, [DynamicStrokeOptions.Solid(Join.Miter, Cap.Butt, Cap.Butt)], PathSoA.from_paths([tight]))
    assert e.value.status == _abi.CR_ERR_CURVE_STEPS_CAPACITY
    rnd.close()


def test_fine_tangent_angle_steps_beyond_the_parameter_buffer(cr, oracle):
    """UniformTangentAngle(0.001) puts thousands of samples into one inflection-free interval (src/curve.rs:228-303 has no limit):
    quadratics stream their samples, cubics sort up to 256 parameters per interval in place and stream longer, ordered runs.
    Every segment kind, an S cubic (two intervals) and a loop cubic, byte for byte against the oracle."""
    w = 0.70710678
    so = dict(width=0.2, offset=0.05, miter_clip=2.0, closed=False, curve_approximation=CurveApproximation.UniformTangentAngle(0.001))
    paths = [
        stroke_path([("Q", [[1, 2], [2, 0]])], [0, 0], **so),
        stroke_path([("RQ", (w, [[1, 1], [0, 1]]))], [1, 0], **so),
        stroke_path([("C", [[3, -1], [4, 2], [5, 0]])], [0, 0], **so),
        stroke_path([("C", [[3, 2], [-1, 2], [2, 0]])], [0, 0], **so),
        stroke_path([("RC", ([1, 0.6, 1.7, 1], [[0.5, 1.5], [2.5, 1.2], [3, 0]]))], [0, 0], **so),
    ]
    dso = [DynamicStrokeOptions.Solid(Join.Round, Cap.Round, Cap.Square)]
    soa = PathSoA.from_paths(paths)
    rnd = cr.Renderer()
    shape = cr.Shape.from_paths(rnd, dso, soa)
    ref = oracle.shape_from_paths(dso, soa)
    assert ref.vertex_offsets[0] // 20 > 5 * 2 * 1000, "the scene is meant to exceed the 256-parameter buffer by far"
    assert_shape_equal(oracle, shape, ref, "fine steps")
    shape.close()
    rnd.close()


def test_device_resident_inputs_match_host_inputs(cr, oracle):
    """CR_MEM_DEVICE inputs (zero-copy) give the same bytes as CR_MEM_HOST inputs."""
    import torch
    scene = scenes.mixed_fills(200, rational=True)
    rnd = cr.Renderer()
    host = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
    tensors = [torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).cuda() for a in scene.paths.arrays()]
    ptrs = [t.data_ptr() if t.numel() else 0 for t in tensors]
    torch.cuda.synchronize()
    dev = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin, memory_space=_abi.CR_MEM_DEVICE, pointers=ptrs)
    for i in range(scene.n_shapes):
        assert np.array_equal(host[i].vertex_buffer(), dev[i].vertex_buffer())
        assert np.array_equal(host[i].index_buffer(), dev[i].index_buffer())
    host.close()
    dev.close()
    rnd.close()


# ---------------------------------------------------------------------------------------------------- raster
@pytest.mark.parametrize("winding_bits", [1, 4], ids=["even_odd", "non_zero"])
def test_config2_fill_rules(cr, oracle, winding_bits):
    """BASELINE config 2 (reduced count, same generator): mixed filled paths, even-odd (1 winding bit) and non-zero."""
    scene = scenes.mixed_fills(400, extent=(640, 360), size=(8.0, 90.0))
    cfg = cr.Configuration(winding_counter_bits=winding_bits, clip_nesting_counter_bits=4)
    got, want = render_both(cr, oracle, scene, cfg)
    assert want[2] > 10000
    assert_frames_equal(got, want)


def test_all_segment_kinds_translucent(cr, oracle):
    scene = scenes.mixed_fills(300, extent=(512, 384), size=(8.0, 100.0), rational=True, seed=99)
    scene.colors[:, 3] = np.linspace(0.2, 1.0, scene.n_shapes, dtype=np.float32)
    got, want = render_both(cr, oracle, scene)
    assert_frames_equal(got, want)


def test_glyph_scene(cr, oracle):
    scene = scenes.glyph_like_fills(1200, extent=(640, 240), glyphs_per_shape=60)
    got, want = render_both(cr, oracle, scene)
    assert want[2] > 5000
    assert_frames_equal(got, want)


def test_text_scene(cr, oracle):
    """BASELINE config 3 generator at reduced size: OpenSans text through the text front-end (src/text.rs)."""
    scene = scenes.text_glyphs(3000, extent=(1024, 256), chars_per_line=160, glyphs_per_shape=80)
    rnd = cr.Renderer()
    batch = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
    for i, ref in enumerate(oracle_shapes(oracle, scene)):
        assert_shape_equal(oracle, batch[i], ref, f"text shape {i}")
    batch.close()
    rnd.close()
    got, want = render_both(cr, oracle, scene)
    assert want[2] > 15000
    assert_frames_equal(got, want)


@pytest.mark.parametrize("dashed", [False, True], ids=["solid", "dashed"])
@pytest.mark.parametrize("join", [Join.Miter, Join.Bevel, Join.Round])
def test_strokes_joins_caps_dashes(cr, oracle, join, dashed):
    scene = scenes.closed_cubic_strokes(60, extent=(512, 384), pixels_per_unit=24.0)
    scene.paths.stroke_options["width"] *= 2.0
    if dashed:
        scene.dynamic_stroke_options = [DynamicStrokeOptions.Dashed(join, [DashInterval(1.0, 2.0, Cap.Round, Cap.Out), DashInterval(3.5, 4.0, Cap.In, Cap.Square)], 0.3)]
    else:
        scene.dynamic_stroke_options = [DynamicStrokeOptions.Solid(join, Cap.Round, Cap.Square)]
    got, want = render_both(cr, oracle, scene)
    assert want[2] > 2000
    assert_frames_equal(got, want)


@pytest.mark.parametrize("caps", [(Cap.Square, Cap.Round), (Cap.Out, Cap.In), (Cap.Right, Cap.Left), (Cap.Butt, Cap.Butt)], ids=lambda c: f"{c[0].name}_{c[1].name}")
def test_open_strokes_caps(cr, oracle, caps):
    scene = scenes.dashed_rational_strokes(150, extent=(512, 384), paths_per_shape=10, pixels_per_unit=4.0)
    scene.dynamic_stroke_options = [DynamicStrokeOptions.Solid(Join.Round, caps[0], caps[1])]
    got, want = render_both(cr, oracle, scene)
    assert want[2] > 1000
    assert_frames_equal(got, want)


def test_config5_kind_dashed_round(cr, oracle):
    """BASELINE config 5 (reduced count, same generator): dashed open rational-cubic strokes, round joins and caps."""
    scene = scenes.dashed_rational_strokes(400, extent=(768, 432), paths_per_shape=40, pixels_per_unit=5.0)
    got, want = render_both(cr, oracle, scene)
    assert want[2] > 1000
    assert_frames_equal(got, want)


@pytest.mark.parametrize("maker", [
    lambda: scenes.mixed_fills(250, extent=(400, 300), size=(8.0, 80.0), rational=True, seed=41),
    lambda: scenes.glyph_like_fills(900, extent=(512, 200), glyphs_per_shape=60),
    lambda: scenes.closed_cubic_strokes(50, extent=(400, 300), pixels_per_unit=20.0),
    lambda: scenes.dashed_rational_strokes(200, extent=(400, 300), paths_per_shape=20, pixels_per_unit=4.0),
], ids=["all_kinds", "glyphs", "cubic_strokes", "dashed_strokes"])
def test_msaa4_matches_oracle(cr, oracle, maker):
    """msaa_sample_count = 4 (the reference demo's setting, examples/showcase/main.rs:11): four stencil bytes and four
    colours per pixel at the WebGPU standard sample positions, shaded per sample (src/shaders.wgsl:35)."""
    scene = maker()
    scene.colors[:, 3] = np.linspace(0.3, 1.0, scene.n_shapes, dtype=np.float32)
    got, want = render_both(cr, oracle, scene, cr.Configuration(msaa_sample_count=4))
    assert got[0].shape == (scene.height, scene.width, 4, 4) and want[2] > 4000
    assert_frames_equal(got, want)
    # the samples of an edge pixel differ (that is the point of multisampling)
    alpha = got[0][..., 3]
    assert np.any(alpha.max(axis=2) != alpha.min(axis=2))


def test_msaa4_clip_and_opacity(cr, oracle):
    scene = scenes.mixed_fills(4, extent=(200, 160), size=(30.0, 70.0), seed=17)
    scene.origins[:] = np.array([[2.5, 2.0], [2.8, 2.2], [2.2, 1.8], [2.6, 2.4]])
    config = cr.Configuration(msaa_sample_count=4, alpha_layer_count=1)
    colors = scene.colors.copy()
    colors[:, 3] = [1.0, 0.6, 0.5, 0.7]
    S, CLIP, UNCLIP, COLOR, SAVE, SCALE, RESTORE = range(7)
    script = [(0, S, 0), (0, CLIP, 1), (1, S, 1), (1, COLOR, 1), (2, S, 1), (2, SAVE, 1), (2, SCALE, 1), (3, S, 1), (3, COLOR, 1),
              (2, S, 1), (2, RESTORE, 1), (0, S, 0), (0, UNCLIP, 0)]
    rnd = cr.Renderer(config)
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
    refs = oracle_shapes(oracle, scene)
    rp = rnd.begin_render_pass()
    rp.set_instances(scene.transforms(), colors)
    ocmds = []
    for shape, op, depth in script:
        rp.set_clip_depth(depth)
        batch[shape].render(rp, range(shape, shape + 1), cr.RenderOperation(op))
        ocmds.append((shape, shape, shape + 1, op, depth, 0, 0))
    rp.submit()
    color, stencil, layer = rnd.read_color(), rnd.read_stencil(), rnd.read_alpha_layer(0)
    ref_color, ref_stencil, ref_layers, _ = oracle.render(config.to_c(), scene.width, scene.height, refs, ocmds, scene.transforms(), colors)
    assert np.array_equal(stencil, ref_stencil) and np.array_equal(color.view(np.uint32), ref_color.view(np.uint32))
    assert np.array_equal(layer.view(np.uint32), ref_layers[0].view(np.uint32))
    batch.close()
    rnd.close()


def test_instancing_and_perspective(cr, oracle):
    """One shape drawn with several instance matrices, including a perspective one (w != 1) and a culled back face."""
    scene = scenes.mixed_fills(12, extent=(384, 256), size=(20.0, 60.0), rational=True, paths_per_shape=12, seed=3)
    base = scene.transforms()[0].reshape(4, 4).T.astype(np.float64)   # columns -> matrix
    mats = []
    for k in range(5):
        a = 0.5 * k
        rot = np.array([[np.cos(a), -np.sin(a), 0, 0], [np.sin(a), np.cos(a), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        persp = np.eye(4)
        persp[3, 0] = 0.02 * k
        persp[3, 1] = -0.015 * k
        shift = np.eye(4)
        shift[0, 3] = 0.8 * k - 1.5
        shift[1, 3] = 0.4 * k - 0.6
        mats.append((persp @ base @ shift @ rot).T.reshape(16))
    mirror = np.diag([-1.0, 1.0, 1.0, 1.0])
    mats.append((base @ mirror).T.reshape(16))                           # flips the facing of every triangle
    transforms = np.asarray(mats, np.float32)
    colors = np.random.default_rng(1).uniform(0.1, 1.0, (len(mats), 4)).astype(np.float32)
    cmds = []
    for i in range(len(mats)):
        cmds += [(0, i, i + 1, 0), (0, i, i + 1, 3)]
    for cull in (cr.CullMode.Off, cr.CullMode.Back, cr.CullMode.Front):
        got, want = render_both(cr, oracle, scene, cr.Configuration(cull_mode=cull), cmds, transforms, colors)
        assert_frames_equal(got, want)
    # one instanced Stencil over all instances, then one instanced Color (instance ranges, src/renderer.rs:271)
    got, want = render_both(cr, oracle, scene, None, [(0, 0, len(mats), 0), (0, 0, len(mats), 3)], transforms, colors)
    assert_frames_equal(got, want)


def test_clip_and_opacity_protocols(cr, oracle):
    """Nested clipping (Stencil -> Clip -> children -> UnClip) and group opacity (Save/Scale/RestoreAlphaContext),
    src/renderer.rs:253-266, with the pass state (clip depth, alpha layer) changing between draws."""
    scene = scenes.mixed_fills(6, extent=(320, 240), size=(40.0, 90.0), seed=11)
    scene.origins[:] = np.array([[4.0, 3.0], [4.5, 3.2], [3.6, 2.8], [4.2, 3.5], [3.9, 2.6], [4.4, 3.1]])
    config = cr.Configuration(alpha_layer_count=2)
    transforms, colors = scene.transforms(), scene.colors.copy()
    colors[:, 3] = [1.0, 0.6, 0.5, 0.7, 0.4, 0.8]
    S, CLIP, UNCLIP, COLOR, SAVE, SCALE, RESTORE = range(7)
    # (shape, op, clip_depth, save_layer, restore_layer)
    script = [
        (0, S, 0, 0, 0), (0, CLIP, 1, 0, 0),                       # clip to shape 0
        (1, S, 1, 0, 0), (1, COLOR, 1, 0, 0),                      # child inside the clip
        (2, S, 1, 0, 0), (2, CLIP, 2, 0, 0),                       # nested clip
        (3, S, 2, 0, 0), (3, SAVE, 2, 0, 0), (3, SCALE, 2, 0, 0),  # opacity group over shape 3's area
        (4, S, 2, 0, 0), (4, COLOR, 2, 0, 0),
        (3, S, 2, 0, 0), (3, RESTORE, 2, 0, 0),
        (2, S, 1, 0, 0), (2, UNCLIP, 1, 0, 0),                     # leave the nested clip
        (5, S, 1, 0, 0), (5, COLOR, 1, 0, 0),
        (0, S, 0, 0, 0), (0, UNCLIP, 0, 0, 0),
    ]
    rnd = cr.Renderer(config)
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
    refs = oracle_shapes(oracle, scene)
    rp = rnd.begin_render_pass()
    rp.set_instances(transforms, colors)
    ocmds = []
    for shape, op, depth, save, restore in script:
        rp.set_clip_depth(depth)
        rp.save_alpha_context(save)
        rp.restore_alpha_context(restore)
        batch[shape].render(rp, range(shape, shape + 1), cr.RenderOperation(op))
        ocmds.append((shape, shape, shape + 1, op, depth, save, restore))
    rp.submit()
    color, stencil, layer = rnd.read_color(), rnd.read_stencil(), rnd.read_alpha_layer(0)
    ref_color, ref_stencil, ref_layers, _ = oracle.render(config.to_c(), scene.width, scene.height, refs, ocmds, transforms, colors)
    assert np.array_equal(stencil, ref_stencil)
    assert np.array_equal(color.view(np.uint32), ref_color.view(np.uint32))
    assert np.array_equal(layer.view(np.uint32), ref_layers[0].view(np.uint32))
    assert (stencil == 0).all(), "clip and winding bits are back to zero after the matching UnClip"
    batch.close()
    rnd.close()


@pytest.mark.parametrize("samples", [1, 4], ids=["1x", "msaa4"])
def test_config4_tiger_like_nested_clips_and_opacity_groups(cr, oracle, samples):
    """BASELINE config 4 at reduced size: placed copies of a 24-Shape group authored with the path constructors
    (src/path.rs:639-815: rounded rectangles, ellipses, arc wedges as rational quadratics and, degree-elevated, as rational
    cubics), every copy running three nested clips and two nested opacity groups (src/renderer.rs:253-266), copies
    overlapping each other, rotated and scaled by their instance matrices."""
    scene = scenes.tiger_like(14, extent=(640, 400), instance_px=(120.0, 330.0))
    config = cr.Configuration(alpha_layer_count=2, msaa_sample_count=samples)
    rnd = cr.Renderer(config)
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
    refs = oracle_shapes(oracle, scene)
    for i, ref in enumerate(refs):
        assert_shape_equal(oracle, batch[i], ref, f"group shape {i}")
    rp = rnd.begin_render_pass()
    rp.set_instances(scene.transforms, scene.colors)
    scene.record(rp, batch, one_call=(samples == 4))   # 1x: the individual state calls; 4x: cr_pass_render_script
    rp.submit()
    color, stencil, covered = rnd.read_color(), rnd.read_stencil(), int(rnd.stats().covered_samples)
    layers = [rnd.read_alpha_layer(k) for k in range(2)]
    ref_color, ref_stencil, ref_layers, ref_covered = oracle.render(config.to_c(), scene.width, scene.height, refs, scene.oracle_commands(),
                                                                    scene.transforms, scene.colors, threads=4)
    assert np.array_equal(stencil, ref_stencil)
    assert np.array_equal(color.view(np.uint32), ref_color.view(np.uint32))
    for k in range(2):
        assert np.array_equal(layers[k].view(np.uint32), ref_layers[k].view(np.uint32))
    assert covered == ref_covered and covered > 20000
    # every copy leaves the clip bits at zero; winding bits only survive on the few boundary pixels that a Shape's own hull
    # misses (the reference's hull drops turns <= 1e-4, src/convex_hull.rs:16)
    assert (stencil >> 4 == 0).all() and int((stencil != 0).sum()) <= 8 * samples
    batch.close()
    rnd.close()


def test_constructed_conics_cover_their_area(cr, oracle):
    """A circle, an ellipse and a rounded rectangle from the constructors (rational quadratics) cover the pixels of the
    exact shape: the implicit test u^2 - vw <= 0 (src/shaders.wgsl:250-257) on constructor output, independent of the
    oracle. (Degree-elevated to rational cubics the same outlines are degenerate cubics — every inflection coefficient
    vanishes — for which the reference's Loop-Blinn classification (src/fill.rs:34-68) has no case: GPU and oracle still
    agree bit for bit, test_config4 covers that, but ~1 % boundary pixels differ from the exact conic.)"""
    ppu, w, h = 40.0, 480, 360
    for cubic in (False,):
        shapes = [Path.from_circle([3.0, 3.0], 2.0), Path.from_ellipse([8.5, 3.0], [2.5, 1.5]), Path.from_rounded_rect([5.0, 7.0], [3.0, 1.2], 0.6)]
        if cubic:
            for p in shapes:
                p.convert_quadratic_curves_to_cubic_curves()
        soa = PathSoA.from_paths(shapes)
        rnd = cr.Renderer()
        rnd.resize_internal_buffers(w, h)
        batch = cr.ShapeBatch(rnd, [], soa, np.arange(4, dtype=np.uint32))
        m = cr.orthographic_transform(w / ppu, h / ppu)
        colors = np.array([[1, 0, 0, 1], [0, 1, 0, 1], [0, 0, 1, 1]], np.float32)
        rp = rnd.begin_render_pass()
        rp.set_instances(np.tile(m, (3, 1)), colors)
        rp.render_batch(batch, scenes.stencil_cover_commands(3))
        rp.submit()
        color = rnd.read_color().reshape(h, w, 4)
        ys, xs = np.mgrid[0:h, 0:w]
        x, y = (xs + 0.5) / ppu, (ys + 0.5) / ppu
        inside_circle = (x - 3.0) ** 2 + (y - 3.0) ** 2 < 4.0
        inside_ellipse = ((x - 8.5) / 2.5) ** 2 + ((y - 3.0) / 1.5) ** 2 < 1.0
        dx, dy = np.maximum(np.abs(x - 5.0) - 2.4, 0.0), np.maximum(np.abs(y - 7.0) - 0.6, 0.0)
        inside_rrect = dx * dx + dy * dy < 0.36
        for channel, inside in ((0, inside_circle), (1, inside_ellipse), (2, inside_rrect)):
            got = color[:, :, channel] > 0.5
            wrong = int((got != inside).sum())
            assert wrong <= 6, f"cubic={cubic} channel {channel}: {wrong} pixels differ from the exact shape"   # centres within 1e-4 px of the outline
            assert int(got.sum()) > 1000
        batch.close()
        rnd.close()


def test_load_op_keeps_previous_pass(cr, oracle):
    """A second pass with LoadOp::Load composites over the first; two submits == one submit of both command lists."""
    scene = scenes.mixed_fills(80, extent=(320, 200), size=(10.0, 60.0), seed=21)
    scene.colors[:, 3] = 0.5
    cmds = scenes.stencil_cover_commands(scene.n_shapes)
    rnd = cr.Renderer()
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
    half = len(cmds) // 2
    for k, part in enumerate((cmds[:half], cmds[half:])):
        rp = rnd.begin_render_pass(clear_color=(k == 0), clear_stencil=(k == 0))
        rp.set_instances(scene.transforms(), scene.colors)
        rp.render_batch(batch, part)
        rp.submit()
    two = (rnd.read_color(), rnd.read_stencil())
    rp = rnd.begin_render_pass()
    rp.set_instances(scene.transforms(), scene.colors)
    rp.render_batch(batch, cmds)
    rp.submit()
    one = (rnd.read_color(), rnd.read_stencil())
    assert np.array_equal(one[1], two[1]) and np.array_equal(one[0].view(np.uint32), two[0].view(np.uint32))
    batch.close()
    rnd.close()


def test_clear_belongs_to_the_pass(cr, oracle):
    """LoadOp::Clear executes with the pass (wgpu semantics): a pass dropped without submit clears nothing; an empty pass that
    is submitted clears; a pass with draws clears every tile it does not touch as well (the tile kernel writes them)."""
    scene = scenes.mixed_fills(12, extent=(200, 120), size=(10.0, 40.0), seed=3)
    cmds = scenes.stencil_cover_commands(scene.n_shapes)
    rnd = cr.Renderer()
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)

    def draw(commands, **clear):
        rp = rnd.begin_render_pass(**clear)
        rp.set_instances(scene.transforms(), scene.colors)
        rp.render_batch(batch, commands)
        rp.submit()
        return rnd.read_color(), rnd.read_stencil()

    full_color, full_stencil = draw(cmds)
    assert np.abs(full_color).sum() > 0
    rp = rnd.begin_render_pass()          # clear requested ...
    rp.close()                            # ... but the pass is dropped: nothing happens
    assert np.array_equal(rnd.read_color().view(np.uint32), full_color.view(np.uint32))
    first_color, first_stencil = draw(cmds[:2])            # one shape only, with clear: the other shapes' tiles are cleared too
    rp = rnd.begin_render_pass()
    rp.submit()                                            # empty pass, submitted: clears
    assert not rnd.read_color().any() and not rnd.read_stencil().any()
    again_color, again_stencil = draw(cmds[:2])
    assert np.array_equal(again_color.view(np.uint32), first_color.view(np.uint32)) and np.array_equal(again_stencil, first_stencil)
    refs = oracle_shapes(oracle, scene)
    ref_color, ref_stencil, _, _ = oracle.render(rnd.config.to_c(), scene.width, scene.height, refs, [(0, 0, 1, 0, 0, 0, 0), (0, 0, 1, 3, 0, 0, 0)],
                                                 scene.transforms(), scene.colors)
    assert np.array_equal(first_color.view(np.uint32), ref_color.view(np.uint32)) and np.array_equal(first_stencil, ref_stencil)
    batch.close()
    rnd.close()


# -------------------------------------------------------------------- size-independent properties at full size
def test_full_size_config3_text_matches_oracle(cr, oracle):
    """BASELINE config 3 at full size through the text front-end (the benchmark workload): 100k OpenSans glyph instances,
    143.6k contour paths, 1.6M segments, 3840x2160 — tessellation of sampled Shapes and the whole frame bit-exact."""
    scene = scenes.text_glyphs(100000)
    assert scene.paths.n_paths == 143596 and scene.paths.n_segments == 1599370 and scene.n_shapes == 625
    rnd = cr.Renderer()
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
    cmds = scenes.stencil_cover_commands(scene.n_shapes)
    rp = rnd.begin_render_pass()
    rp.set_instances(scene.transforms(), scene.colors)
    rp.render_batch(batch, cmds)
    rp.submit()
    got = (rnd.read_color(), rnd.read_stencil(), int(rnd.stats().covered_samples))
    refs = oracle_shapes(oracle, scene)
    for i in (0, 1, scene.n_shapes // 2, scene.n_shapes - 1):
        assert_shape_equal(oracle, batch[i], refs[i], f"shape {i}")
    ocmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in cmds]
    ref_color, ref_stencil, _, ref_covered = oracle.render(rnd.config.to_c(), scene.width, scene.height, refs, ocmds, scene.transforms(), scene.colors,
                                                           threads=oracle.max_threads())
    assert ref_covered > 500000
    assert_frames_equal(got, (ref_color, ref_stencil, ref_covered))
    batch.close()
    rnd.close()


def test_full_size_config3_matches_oracle_and_properties(cr, oracle):
    """A synthetic stand-in of config 3 at full size (100k glyph-sized contour groups, 144k paths, 3840x2160) with model
    coordinates of +-200 units, where f32 noise exceeds the reference's tolerances: bit-exact against the oracle, plus
    size-independent properties: rendering twice gives identical bits (the raster has no order-dependent atomics),
    opaque covers leave alpha in {0, 1}, the cover zeroes the winding bits (src/renderer.rs:747-752).

    The last property holds up to the reference's own hull tolerance: convex_hull::andrew pops points whose turn area is
    <= 1e-4 (src/convex_hull.rs:16) in f32 on absolute coordinates, so a hull may miss a boundary pixel of its own shape
    and the winding bits stay set there. The oracle reproduces exactly the same pixels (17 for this seed)."""
    scene = scenes.glyph_like_fills(100000)
    rnd = cr.Renderer()
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
    cmds = scenes.stencil_cover_commands(scene.n_shapes)
    frames = []
    for _ in range(2):
        rp = rnd.begin_render_pass()
        rp.set_instances(scene.transforms(), scene.colors)
        rp.render_batch(batch, cmds)
        rp.submit()
        frames.append((rnd.read_color(), rnd.read_stencil(), int(rnd.stats().covered_samples)))
    (c0, s0, n0), (c1, s1, n1) = frames
    assert np.array_equal(c0.view(np.uint32), c1.view(np.uint32)) and np.array_equal(s0, s1) and n0 == n1
    assert int((s0 != 0).sum()) < 100
    alpha = c0[..., 3]
    assert set(np.unique(alpha)) <= {0.0, 1.0}
    assert n0 >= int((alpha == 1.0).sum()) > 1000000
    refs = oracle_shapes(oracle, scene)
    for i in (0, 1, scene.n_shapes // 2, scene.n_shapes - 1):
        assert_shape_equal(oracle, batch[i], refs[i], f"shape {i}")
    ocmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in cmds]
    ref_color, ref_stencil, _, ref_covered = oracle.render(rnd.config.to_c(), scene.width, scene.height, refs, ocmds, scene.transforms(), scene.colors,
                                                           threads=oracle.max_threads())
    assert_frames_equal((c0, s0, n0), (ref_color, ref_stencil, ref_covered))
    batch.close()
    rnd.close()


def test_tile_sharded_target_matches_single_gpu():
    """One render target spanning the GPUs of the box (BASELINE config 4): every rank's copy of the frame is bit-identical
    to the single-GPU frame. Needs >= 2 GPUs (`gpurun --gpus 2`); runs tests/multi_gpu/tile_sharding_check.py under torchrun."""
    import json
    import os
    import subprocess
    import sys
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29671",
           os.path.join(root, "tests", "multi_gpu", "tile_sharding_check.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert proc.returncode == 0 and lines, proc.stdout[-2000:] + proc.stderr[-2000:]
    result = json.loads(lines[-1])
    assert result["ok"] and result["identical_and_empty_pass_cleared_on_every_rank"] and result["n_gpus"] == n


def test_order_sharded_target_matches_oracle():
    """One render target composed from draw-order slices over the GPUs of the box (BASELINE config 5): every rank's copy of the
    frame is bit-identical to the CPU oracle's frame of the whole scene. Needs >= 2 GPUs (`gpurun --gpus 2`); runs
    tests/multi_gpu/order_sharding_check.py under torchrun."""
    import json
    import os
    import subprocess
    import sys
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29673",
           os.path.join(root, "tests", "multi_gpu", "order_sharding_check.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert proc.returncode == 0 and lines, proc.stdout[-2000:] + proc.stderr[-2000:]
    result = json.loads(lines[-1])
    assert result["ok"] and result["n_gpus"] == n and all(v["identical_on_every_rank"] for v in result["scenes"].values())
