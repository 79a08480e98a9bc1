"""Depth test / depth write of the colour cover (Configuration::depth_compare, depth_write_enabled, src/renderer.rs:388-390,
743-745; depth-fail keeps the stencil value, :442; clip z from src/shaders.wgsl:72) and the 8-bit colour formats
(Bgra8Unorm surface of examples/application_framework.rs:175, R8Unorm alpha layers of src/renderer.rs:783,898).

CPU part: the oracle against an independent expectation (nearest instance wins whatever the draw order).
GPU part: the CUDA path through the C-ABI against the oracle, bit for bit."""
import numpy as np
import pytest

from contrast_renderer_b200 import scenes, utils
from contrast_renderer_b200.path import Path, PathSoA
from contrast_renderer_b200.renderer import ColorFormat, CompareFunction, Configuration

W, H = 256, 192


def placed_in_3d():
    """Two unit squares of one Shape placed like the showcase places its instances (perspective_projection x a rigid
    placement, examples/showcase/main.rs:163-201): instance 0 nearer to the eye than instance 1, overlapping on screen, the
    second one tilted so that its depth varies across the shape."""
    proj = np.asarray(utils.perspective_projection(np.pi * 0.5, W / H, 1.0, 1000.0), np.float64).reshape(4, 4).T   # columns -> matrix

    def place(tx, ty, tz, tilt):
        c, s = np.cos(tilt), np.sin(tilt)
        rot_y = np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1.0]])
        move = np.eye(4)
        move[:3, 3] = (tx, ty, tz)
        return (proj @ move @ rot_y).T.reshape(16)

    transforms = np.asarray([place(-0.4, 0.1, 3.0, 0.0), place(0.5, -0.2, 4.0, 0.9)], np.float32)
    colors = np.asarray([[1.0, 0.2, 0.1, 1.0], [0.1, 0.3, 1.0, 1.0]], np.float32)
    square = Path.from_rect([0.0, 0.0], [1.2, 1.2])
    return PathSoA.from_paths([square]), transforms, colors


def commands(order):
    out = []
    for i in order:
        out += [(0, i, i + 1, 0, 0, 0, 0), (0, i, i + 1, 3, 0, 0, 0)]
    return out


def oracle_frame(oracle, cfg, soa, transforms, colors, order, samples=1):
    shape = oracle.shape_from_paths([], soa)
    depth = np.ones((H, W, samples), np.float32) if cfg.has_depth else None
    color, stencil, _, covered = oracle.render(cfg.to_c(), W, H, [shape], commands(order), transforms, colors, depth=depth)
    return color, stencil, depth, covered


def test_depth_test_makes_the_image_independent_of_draw_order(oracle):
    soa, transforms, colors = placed_in_3d()
    cfg = Configuration(depth_compare=CompareFunction.LessEqual, depth_write_enabled=True)
    back_to_front = oracle_frame(oracle, cfg, soa, transforms, colors, [1, 0])
    front_to_back = oracle_frame(oracle, cfg, soa, transforms, colors, [0, 1])
    assert np.array_equal(back_to_front[0], front_to_back[0]) and np.array_equal(back_to_front[2], front_to_back[2])
    color, _, depth, _ = front_to_back
    near, far = (color[..., 0, 0] == np.float32(1.0)), (color[..., 0, 2] == np.float32(1.0))
    assert near.sum() > 3000 and far.sum() > 800
    # the near square is 3 units from the eye: z/w = far/(far-near) * (1 - near/w) with near = 1, far = 1000 (src/utils.rs:171-181)
    assert np.allclose(depth[near, 0], (1000.0 / 999.0) * (1.0 - 1.0 / 3.0), atol=1e-6)
    assert (depth[far, 0] > depth[near, 0].max()).all() and np.ptp(depth[far, 0]) > 0.01   # tilted: depth varies across the shape
    assert (depth[~near & ~far, 0] == 1.0).all()
    # without the depth test the draw order decides (the overlap takes the colour of the later draw)
    plain = Configuration()
    a = oracle_frame(oracle, plain, soa, transforms, colors, [1, 0])[0]
    b = oracle_frame(oracle, plain, soa, transforms, colors, [0, 1])[0]
    overlap = (a[..., 0, 0] == 1.0) & (b[..., 0, 2] == 1.0)
    assert overlap.sum() > 500 and np.array_equal(a[overlap], color[overlap]) and not np.array_equal(b[overlap], color[overlap])


def test_depth_failure_keeps_the_stencil_value(oracle):
    """src/renderer.rs:442: depth_fail_op = Keep. Front to back under Less: the far instance's cover fails the depth test inside
    the overlap, so its winding bits stay there (and are zeroed everywhere else it covers)."""
    soa, transforms, colors = placed_in_3d()
    cfg = Configuration(depth_compare=CompareFunction.Less, depth_write_enabled=True)
    color, stencil, depth, covered = oracle_frame(oracle, cfg, soa, transforms, colors, [0, 1])
    residue = stencil[..., 0] != 0
    near = color[..., 0, 0] == np.float32(1.0)
    assert residue.sum() > 500 and near[residue].all()
    assert covered == int(near.sum()) + int((color[..., 0, 2] == np.float32(1.0)).sum())


def test_unorm8_targets_quantise_every_blend(oracle):
    """A translucent cover over an opaque one in Rgba8Unorm: dst = unorm8(src + dst * (1 - a)), one rounding per blend."""
    soa, transforms, colors = placed_in_3d()
    colors = colors.copy()
    colors[1, 3] = 0.4
    cfg = Configuration(color_format=ColorFormat.Rgba8Unorm)
    color = oracle_frame(oracle, cfg, soa, transforms, colors, [0, 1])[0]
    assert np.array_equal(np.rint(color * 255.0) / np.float32(255.0), color)          # every stored value is k / 255
    q = lambda x: np.floor(np.clip(np.float32(x), 0, 1) * np.float32(255.0) + np.float32(0.5)) / np.float32(255.0)
    first = q(colors[0, :3] * colors[0, 3])
    a = np.float32(0.4)
    second = q(colors[1, :3] * a + first * (np.float32(1.0) - a))
    both = np.all(color[..., 0, :3] == second, axis=-1)
    assert both.sum() > 500


# ------------------------------------------------------------------------------------------------------------ GPU parity
def gpu_frame(cr, cfg, soa, transforms, colors, order):
    rnd = cr.Renderer(cfg)
    rnd.resize_internal_buffers(W, H)
    shape = cr.Shape.from_paths(rnd, [], soa)
    rp = rnd.begin_render_pass()
    rp.set_instances(transforms, colors)
    for i in order:
        shape.render(rp, range(i, i + 1), cr.RenderOperation.Stencil)
        shape.render(rp, range(i, i + 1), cr.RenderOperation.Color)
    rp.submit()
    out = (rnd.read_color(), rnd.read_stencil(), rnd.read_depth() if cfg.has_depth else None, int(rnd.stats().covered_samples))
    shape.close()
    rnd.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("samples", [1, 4], ids=["1x", "msaa4"])
@pytest.mark.parametrize("compare", [CompareFunction.LessEqual, CompareFunction.Less, CompareFunction.Greater, CompareFunction.Always])
def test_depth_matches_oracle(cr, oracle, compare, samples):
    soa, transforms, colors = placed_in_3d()
    cfg = Configuration(msaa_sample_count=samples, depth_compare=compare, depth_write_enabled=True)
    for order in ([0, 1], [1, 0], [0, 1, 0]):
        got = gpu_frame(cr, cfg, soa, transforms, colors, order)
        want = oracle_frame(oracle, cfg, soa, transforms, colors, order, samples)
        assert np.array_equal(got[1], want[1]), "stencil"
        assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32)), "colour"
        assert np.array_equal(got[2].view(np.uint32), want[2].view(np.uint32)), "depth"
        assert got[3] == want[3]


@pytest.mark.gpu
def test_depth_attachment_persists_across_passes(cr, oracle):
    """LoadOp::Load of the depth aspect: a second pass that does not clear depth is tested against the first pass's depths."""
    soa, transforms, colors = placed_in_3d()
    cfg = Configuration(depth_compare=CompareFunction.LessEqual, depth_write_enabled=True)
    rnd = cr.Renderer(cfg)
    rnd.resize_internal_buffers(W, H)
    shape = cr.Shape.from_paths(rnd, [], soa)
    for i, clear in ((0, True), (1, False)):
        rp = rnd.begin_render_pass(clear_color=clear, clear_stencil=clear)
        rp.set_instances(transforms, colors)
        shape.render(rp, range(i, i + 1), cr.RenderOperation.Stencil)
        shape.render(rp, range(i, i + 1), cr.RenderOperation.Color)
        rp.submit()
    got = (rnd.read_color(), rnd.read_stencil(), rnd.read_depth())
    want = oracle_frame(oracle, cfg, soa, transforms, colors, [0, 1])
    assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32)) and np.array_equal(got[1], want[1])
    assert np.array_equal(got[2].view(np.uint32), want[2].view(np.uint32))
    with pytest.raises(cr.Error):
        cr.Renderer(Configuration()).read_depth()
    shape.close()
    rnd.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", [ColorFormat.Rgba8Unorm, ColorFormat.Bgra8Unorm], ids=["rgba8", "bgra8"])
@pytest.mark.parametrize("samples", [1, 4], ids=["1x", "msaa4"])
def test_unorm8_config4_matches_oracle(cr, oracle, fmt, samples):
    """The config-4 scene (nested clips, two nested opacity groups, translucent covers) into an 8-bit target with R8 alpha
    layers: colour, stencil and both layers bit for bit against the oracle."""
    scene = scenes.tiger_like(10, extent=(512, 320), instance_px=(120.0, 300.0))
    config = cr.Configuration(alpha_layer_count=2, msaa_sample_count=samples, color_format=fmt)
    rnd = cr.Renderer(config)
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, [], scene.paths, scene.shape_path_begin)
    refs = [oracle.shape_from_paths([], scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1])) for i in range(scene.n_shapes)]
    rp = rnd.begin_render_pass()
    rp.set_instances(scene.transforms, scene.colors)
    scene.record(rp, batch)
    rp.submit()
    color, stencil = rnd.read_color(), rnd.read_stencil()
    layers = [rnd.read_alpha_layer(k) for k in range(2)]
    ref_color, ref_stencil, ref_layers, ref_covered = oracle.render(config.to_c(), scene.width, scene.height, refs, scene.oracle_commands(),
                                                                    scene.transforms, scene.colors, threads=4)
    assert np.array_equal(stencil, ref_stencil)
    assert np.array_equal(color.view(np.uint32), ref_color.view(np.uint32))
    for k in range(2):
        assert np.array_equal(layers[k].view(np.uint32), ref_layers[k].view(np.uint32))
    assert int(rnd.stats().covered_samples) == ref_covered
    # a second pass that LOADS the 8-bit attachments (no clear) continues from the stored texels
    rp = rnd.begin_render_pass(clear_color=False, clear_stencil=False)
    rp.set_instances(scene.transforms, scene.colors)
    scene.record(rp, batch)
    rp.submit()
    again, _, _, _ = oracle.render(config.to_c(), scene.width, scene.height, refs, scene.oracle_commands(), scene.transforms, scene.colors,
                                   color=ref_color, stencil=ref_stencil, alpha_layers=ref_layers, threads=4)
    assert np.array_equal(rnd.read_color().view(np.uint32), again.view(np.uint32))
    batch.close()
    rnd.close()


# ------------------------------------------------------------------------------------------------ frustum clipping
def eye_crossing_transforms(base_columns, n=5):
    """Instance matrices whose projective row makes w change sign inside the shapes (tilted planes that pass the eye), a steep
    one that throws vertices far outside the 2^21-pixel range, and an ordinary one."""
    base = np.asarray(base_columns, np.float64).reshape(4, 4).T
    mats = []
    for k in range(n):
        persp = np.eye(4)
        persp[3, 0] = (-0.9 - 0.35 * k) * (1 if k % 2 == 0 else -1)     # w = 1 + a x_ndc + b y_ndc: zero inside the frame
        persp[3, 1] = 0.25 * k - 0.4
        persp[2, :] = 0.5 * persp[3, :] + np.array([0.05, -0.03, 0.0, 0.0])    # z / w = 0.5 + a term that varies over the shape (depth test)
        mats.append((persp @ base).T.reshape(16))
    far = np.eye(4)
    far[3, 0], far[3, 3] = 1.0, 1.0e-6 + 1.0                                 # nearly singular: some vertices land ~1e6 NDC units away
    far[3, 0] = -0.999999
    mats.append((far @ base).T.reshape(16))
    mats.append(base.T.reshape(16))
    return np.asarray(mats, np.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("samples", [1, 4], ids=["1x", "msaa4"])
def test_frustum_clipped_instances_match_oracle(cr, oracle, samples):
    """Triangles with corners behind the eye plane or out of snapping range are clipped (not dropped), identically on both sides:
    fills of every segment kind (interpolated implicit-curve attributes at the new corners), covers with a depth test (z / w at the
    new corners), culling decided per fan triangle, 1x and 4x."""
    from contrast_renderer_b200 import scenes
    scene = scenes.mixed_fills(10, extent=(320, 240), size=(60.0, 160.0), rational=True, paths_per_shape=5, seed=21)
    transforms = eye_crossing_transforms(scene.transforms()[0])
    colors = np.random.default_rng(5).uniform(0.2, 1.0, (len(transforms), 4)).astype(np.float32)
    colors[:, 3] = [1.0, 0.6, 1.0, 0.8, 1.0, 0.7, 1.0][:len(transforms)]
    cmds = []
    for s in range(scene.n_shapes):
        for i in range(len(transforms)):
            cmds += [(s, i, i + 1, 0), (s, i, i + 1, 3)]
    cull = 1   # CullMode.Front: these matrices mirror the shapes, their hulls face backwards
    for cfg in (Configuration(msaa_sample_count=samples),
                Configuration(msaa_sample_count=samples, depth_compare=CompareFunction.LessEqual, depth_write_enabled=True, cull_mode=cull)):
        rnd = cr.Renderer(cfg)
        rnd.resize_internal_buffers(scene.width, scene.height)
        batch = cr.ShapeBatch(rnd, scene.dynamic_stroke_options, scene.paths, scene.shape_path_begin)
        rp = rnd.begin_render_pass()
        rp.set_instances(transforms, colors)
        rp.render_batch(batch, np.asarray(cmds, np.uint32))
        rp.submit()
        color, stencil, stats = rnd.read_color(), rnd.read_stencil(), rnd.stats()
        refs = [oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
                for i in range(scene.n_shapes)]
        depth = np.ones((scene.height, scene.width, samples), np.float32) if cfg.has_depth else None
        ocmds = [(c[0], c[1], c[2], c[3], 0, 0, 0) for c in cmds]
        ref_color, ref_stencil, _, ref_covered = oracle.render(cfg.to_c(), scene.width, scene.height, refs, ocmds, transforms, colors, depth=depth, threads=4)
        assert ref_covered > 20000, "the clipped instances must actually draw (and survive culling / the depth test)"
        assert np.array_equal(stencil, ref_stencil), f"stencil differs at {np.argwhere(stencil != ref_stencil)[:5]}"
        assert np.array_equal(color.view(np.uint32), ref_color.view(np.uint32))
        assert int(stats.covered_samples) == ref_covered
        if cfg.has_depth:
            assert np.array_equal(rnd.read_depth().view(np.uint32), depth.view(np.uint32))
        batch.close()
        rnd.close()


@pytest.mark.gpu
def test_frustum_clipped_strokes_and_capacity_growth(cr, oracle):
    """Stroke triangles keep their FLAT attributes (path index / cap flags of the original first vertex) through clipping, dashes and
    caps included; the scene clips far more triangles than the renderer's initial clip capacity (1024), so the pass is re-sized
    and re-submitted, twice in a row with the same result (the second pass runs optimistically with the grown capacity)."""
    from contrast_renderer_b200 import scenes
    scene = scenes.dashed_rational_strokes(160, paths_per_shape=20, extent=(320, 240))
    transforms = eye_crossing_transforms(scene.transforms()[0], n=3)
    colors = np.random.default_rng(6).uniform(0.2, 1.0, (len(transforms), 4)).astype(np.float32)
    cmds = []
    for s in range(scene.n_shapes):
        cmds += [(s, 0, len(transforms), 0), (s, 0, len(transforms), 3)]
    cfg = Configuration()
    rnd = cr.Renderer(cfg)
    rnd.resize_internal_buffers(scene.width, scene.height)
    batch = cr.ShapeBatch(rnd, scene.dynamic_stroke_options, scene.paths, scene.shape_path_begin)
    refs = [oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
            for i in range(scene.n_shapes)]
    ocmds = [(c[0], c[1], c[2], c[3], 0, 0, 0) for c in cmds]
    ref_color, ref_stencil, _, ref_covered = oracle.render(cfg.to_c(), scene.width, scene.height, refs, ocmds, transforms, colors, threads=4)
    assert ref_covered > 5000
    for _ in range(2):
        rp = rnd.begin_render_pass()
        rp.set_instances(transforms, colors)
        rp.render_batch(batch, np.asarray(cmds, np.uint32))
        rp.submit()
        color, stencil = rnd.read_color(), rnd.read_stencil()
        assert np.array_equal(stencil, ref_stencil), f"stencil differs at {np.argwhere(stencil != ref_stencil)[:5]}"
        assert np.array_equal(color.view(np.uint32), ref_color.view(np.uint32))
        assert int(rnd.stats().covered_samples) == ref_covered
    batch.close()
    rnd.close()
