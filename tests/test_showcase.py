"""The reference's own worked scene (examples/showcase/main.rs) end to end: "Hello World" through the text front-end inside a
dashed rounded rectangle, ONE Shape drawn as 46 instances placed in 3D by ppga3d motors and a perspective projection, 4x MSAA,
depth LessEqual + depth write, back-face culling, Stencil + Color per instance — built with the host mirrors of the
reference's helpers (`scenes.showcase`), rendered through the C-ABI and compared with the oracle bit for bit; with the camera
turned so far that instances cross the eye plane (frustum clipping) as well."""
import numpy as np
import pytest

from contrast_renderer_b200 import scenes, utils
from contrast_renderer_b200.renderer import CompareFunction, Configuration

W, H = 480, 270


def showcase_config(samples=4):
    return Configuration(msaa_sample_count=samples, depth_compare=CompareFunction.LessEqual, depth_write_enabled=True, cull_mode=2)


def commands(n):
    out = []
    for i in range(n):
        out += [(0, i, i + 1, 0, 0, 0, 0), (0, i, i + 1, 3, 0, 0, 0)]
    return out


def test_showcase_scene_is_the_demo(oracle, tmp_path):
    soa, begin, dso, transforms, colors = scenes.showcase((W, H))
    assert soa.n_paths == 15 and len(transforms) == 46 and list(begin) == [0, 15]
    # instance 0 sits view_distance in front of the eye, the grid 10 further away (main.rs:171,190)
    origin_w = [float((t.reshape(4, 4).T @ np.array([0, 0, 0, 1.0]))[3]) for t in transforms]
    assert np.isclose(origin_w[0], 5.0) and np.allclose(origin_w[1:], 15.0)
    shape = oracle.shape_from_paths(dso, soa)
    depth = np.ones((H, W, 4), np.float32)
    color, stencil, _, covered = oracle.render(showcase_config().to_c(), W, H, [shape], commands(46), transforms, colors, depth=depth, threads=4)
    assert covered > 40000
    # where a grid instance lies behind the nearer instance 0 its colour cover fails the depth test, and a depth failure KEEPS the
    # stencil value (src/renderer.rs:442): the winding residue stays there, exactly as in the reference's demo
    assert stencil.any() and (stencil != 0).mean() < 0.05
    # the grid cells in view show their instance colour; the nearer instance 0 (white) wins where it overlaps them (depth test)
    resolved = color.mean(axis=2)
    assert resolved[..., 3].max() > 0.99 and (resolved[..., :3].max(axis=(0, 1)) > 0.9).all()
    utils.save_png(str(tmp_path / "showcase.png"), resolved)
    assert (tmp_path / "showcase.png").stat().st_size > 4000


@pytest.mark.gpu
@pytest.mark.parametrize("angles", [(0.0, 0.0), (1.1, 0.5), (-2.4, 0.3)], ids=["front", "turned", "from_behind"])
def test_showcase_matches_oracle(cr, oracle, angles):
    soa, begin, dso, transforms, colors = scenes.showcase((W, H), angles, 5.0)
    cfg = showcase_config()
    rnd = cr.Renderer(cfg)
    rnd.resize_internal_buffers(W, H)
    shape = cr.Shape.from_paths(rnd, dso, soa)
    rp = rnd.begin_render_pass()
    rp.set_instances(transforms, colors)
    for i in range(len(transforms)):
        shape.render(rp, range(i, i + 1), cr.RenderOperation.Stencil)
        shape.render(rp, range(i, i + 1), cr.RenderOperation.Color)
    rp.submit()
    color, stencil, depth, covered = rnd.read_color(), rnd.read_stencil(), rnd.read_depth(), int(rnd.stats().covered_samples)
    ref = oracle.shape_from_paths(dso, soa)
    ref_depth = np.ones((H, W, 4), np.float32)
    ref_color, ref_stencil, _, ref_covered = oracle.render(cfg.to_c(), W, H, [ref], commands(len(transforms)), transforms, colors, depth=ref_depth, threads=4)
    assert np.array_equal(shape.vertex_buffer(), ref.vertex_buffer) and np.array_equal(shape.index_buffer(), ref.index_buffer)
    assert np.array_equal(stencil, ref_stencil)
    assert np.array_equal(color.view(np.uint32), ref_color.view(np.uint32))
    assert np.array_equal(depth.view(np.uint32), ref_depth.view(np.uint32))
    assert covered == ref_covered
    if angles == (0.0, 0.0):
        assert covered > 40000
    shape.close()
    rnd.close()
