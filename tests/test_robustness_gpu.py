"""Inputs the reference's API cannot express but a C caller can: inconsistent cr_path_soa cursor tables, and passes whose
(tile, primitive) pair count does not fit the 32-bit scan. Both must come back as CR_ERR_INVALID_ARGUMENT, not as
out-of-bounds device accesses or a corrupt frame."""
import copy

import numpy as np
import pytest

from contrast_renderer_b200 import _abi, scenes
from contrast_renderer_b200.path import Path, PathSoA

pytestmark = pytest.mark.gpu


def _expect_invalid(cr, rnd, soa):
    with pytest.raises(cr.Error) as e:
        cr.ShapeBatch(rnd, [], soa, np.array([0, soa.n_paths], np.uint32))
    assert e.value.status == _abi.CR_ERR_INVALID_ARGUMENT


def test_inconsistent_cursor_tables_are_rejected(cr):
    scene = scenes.mixed_fills(40, extent=(256, 192), rational=True, seed=3)
    rnd = cr.Renderer()
    good = scene.paths
    cr.ShapeBatch(rnd, [], good, scene.shape_path_begin).close()   # the untouched tables are accepted

    bad = copy.deepcopy(good)
    bad.segment_begin[5], bad.segment_begin[6] = good.segment_begin[6], good.segment_begin[5]   # not monotone
    _expect_invalid(cr, rnd, bad)

    bad = copy.deepcopy(good)
    bad.segment_begin[-1] += 3                                                                   # beyond n_segments
    _expect_invalid(cr, rnd, bad)

    bad = copy.deepcopy(good)
    bad.segment_types[int(good.segment_begin[7])] = 9                                            # not a cr_segment_type
    _expect_invalid(cr, rnd, bad)

    bad = copy.deepcopy(good)
    t = int(good.segment_types[int(good.segment_begin[3])])
    bad.type_begin[t, 4:] += 1                                                                   # per-type count disagrees with the type stream
    _expect_invalid(cr, rnd, bad)

    bad = copy.deepcopy(good)
    bad.type_begin[2, 10] = bad.type_begin[2, 11] + 5                                            # a per-type cursor runs backwards
    _expect_invalid(cr, rnd, bad)

    # the renderer is still usable afterwards
    batch = cr.ShapeBatch(rnd, [], good, scene.shape_path_begin)
    assert len(batch) == scene.n_shapes
    batch.close()
    rnd.close()


def test_pair_count_beyond_32_bits_is_an_error(cr):
    """129 600 tiles at 8K x 34 000 instances of a target-filling hull cover = 8.8e9 (tile, primitive) pairs."""
    rnd = cr.Renderer()
    rnd.resize_internal_buffers(7680, 4320)
    shape = cr.Shape.from_paths(rnd, [], PathSoA.from_paths([Path.from_rect([0.0, 0.0], [2.0, 2.0])]))   # covers all of NDC under the identity
    n = 34000
    transforms = np.tile(np.eye(4, dtype=np.float32).reshape(-1), (n, 1))
    colors = np.ones((n, 4), np.float32)
    rp = rnd.begin_render_pass()
    rp.set_instances(transforms, colors)
    shape.render(rp, range(0, n), cr.RenderOperation.Color)
    with pytest.raises(cr.Error) as e:
        rp.submit()
    assert e.value.status == _abi.CR_ERR_INVALID_ARGUMENT and "pairs" in str(e.value)
    # a pass that fits still renders
    rp = rnd.begin_render_pass()
    rp.set_instances(transforms[:1], colors[:1])
    shape.render(rp, range(0, 1), cr.RenderOperation.Stencil)
    shape.render(rp, range(0, 1), cr.RenderOperation.Color)
    rp.submit()
    assert int(rnd.stats().covered_samples) == 7680 * 4320
    shape.close()
    rnd.close()


def _render_scene(cr, rnd, scene, batch=None, memory_space=None, pointers=None):
    kwargs = {}
    if memory_space is not None:
        kwargs = dict(memory_space=memory_space, pointers=pointers)
    batch = cr.ShapeBatch(rnd, scene.dynamic_stroke_options, scene.paths, scene.shape_path_begin, existing=batch, **kwargs)
    rp = rnd.begin_render_pass()
    rp.set_instances(scene.transforms(), scene.colors)
    rp.render_batch(batch, scenes.stencil_cover_commands(scene.n_shapes))
    rp.submit()
    return batch


def _oracle_frame(oracle, rnd, scene):
    refs = [oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
            for i in range(scene.n_shapes)]
    cmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in scenes.stencil_cover_commands(scene.n_shapes)]
    color, stencil, _, covered = oracle.render(rnd.config.to_c(), scene.width, scene.height, refs, cmds, scene.transforms(), scene.colors, threads=4)
    return color, stencil, covered


def test_optimistic_pass_is_resubmitted_when_a_capacity_is_exceeded(cr, oracle):
    """cr_pass_submit sizes the second and later passes of a renderer from what the previous pass needed and only checks the
    device-side totals afterwards. A pass that needs more candidates and more (tile, primitive) pairs than its predecessor must
    still produce the right frame (the tile kernel does not run on the undersized attempt; the pass is submitted again)."""
    small = scenes.mixed_fills(12, extent=(512, 384), size=(10.0, 40.0), seed=5)
    large = scenes.mixed_fills(500, extent=(512, 384), size=(10.0, 120.0), rational=True, seed=6)
    rnd = cr.Renderer()
    rnd.resize_internal_buffers(512, 384)
    for scene in (small, large, small, large):
        batch = _render_scene(cr, rnd, scene)
        color, stencil = rnd.read_color(), rnd.read_stencil()
        ref_color, ref_stencil, ref_covered = _oracle_frame(oracle, rnd, scene)
        assert np.array_equal(stencil, ref_stencil) and np.array_equal(color.view(np.uint32), ref_color.view(np.uint32))
        assert int(rnd.stats().covered_samples) == ref_covered
        batch.close()
    # back-to-back passes without any read in between: each is settled by the next call on the renderer
    batches = []
    for scene in (small, large):
        batches.append(_render_scene(cr, rnd, scene))
    ref_color, ref_stencil, _ = _oracle_frame(oracle, rnd, large)
    assert np.array_equal(rnd.read_stencil(), ref_stencil) and np.array_equal(rnd.read_color().view(np.uint32), ref_color.view(np.uint32))
    for b in batches:
        b.close()
    rnd.close()


def test_optimistic_rebuild_falls_back_when_the_outputs_grow_or_change_kind(cr, oracle):
    """cr_shape_batch_from_paths with `existing`: the emit pass is enqueued into the previous build's arrays before the host
    knows the new sizes. Same numbers of paths / shapes / segments, but (a) a finer curve approximation (more vertices than
    the arrays hold) and (b) cubic segments where the previous build had none (device inputs: the kernels without the cubic
    builder were launched) must both end in the right buffers."""
    import torch
    coarse = scenes.closed_cubic_strokes(60, extent=(512, 384), angle_step=0.4)
    fine = scenes.closed_cubic_strokes(60, extent=(512, 384), angle_step=0.05)
    assert coarse.paths.n_segments == fine.paths.n_segments
    rnd = cr.Renderer()
    rnd.resize_internal_buffers(512, 384)
    batch = None
    for scene in (coarse, fine, coarse):
        batch = _render_scene(cr, rnd, scene, batch)
        for i in range(0, scene.n_shapes, 7):
            ref = oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
            assert np.array_equal(batch[i].vertex_buffer(), ref.vertex_buffer) and np.array_equal(batch[i].index_buffer(), ref.index_buffer)
        ref_color, ref_stencil, _ = _oracle_frame(oracle, rnd, scene)
        assert np.array_equal(rnd.read_stencil(), ref_stencil) and np.array_equal(rnd.read_color().view(np.uint32), ref_color.view(np.uint32))
    batch.close()
    # (b) device-resident inputs: quadratics only, then the same segment count as cubics
    quads = scenes.mixed_fills(80, extent=(512, 384), seed=9, types=(1,))
    cubics = scenes.mixed_fills(80, extent=(512, 384), seed=9, types=(2,))
    assert quads.paths.n_segments == cubics.paths.n_segments and quads.paths.n_paths == cubics.paths.n_paths
    batch = None
    for scene in (quads, cubics, quads):
        arrays = scene.paths.arrays()
        tensors = [torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).cuda() for a in arrays]
        ptrs = [t.data_ptr() if t.numel() else 0 for t in tensors]
        ptrs[9] = 0
        torch.cuda.synchronize()
        batch = _render_scene(cr, rnd, scene, batch, memory_space=_abi.CR_MEM_DEVICE, pointers=ptrs)
        ref_color, ref_stencil, _ = _oracle_frame(oracle, rnd, scene)
        assert np.array_equal(rnd.read_stencil(), ref_stencil) and np.array_equal(rnd.read_color().view(np.uint32), ref_color.view(np.uint32))
    batch.close()
    rnd.close()


def test_frame_pipelining_renders_the_same_frames(cr, oracle):
    """cr_renderer_set_pipelining: rebuilds go to a second stream and to the batch's other set of arrays while the previous
    pass is still being rasterised. A sequence of different frames rendered back to back (no read in between, then one read per
    frame in a second round) must give the oracle's frames, including a frame whose sizes exceed every capacity."""
    frames = [scenes.mixed_fills(60, extent=(512, 384), size=(10.0, 60.0), seed=21, types=(0, 1)),
              scenes.mixed_fills(60, extent=(512, 384), size=(10.0, 60.0), seed=22, types=(0, 1)),
              scenes.mixed_fills(60, extent=(512, 384), size=(20.0, 140.0), seed=23, types=(0, 1)),   # larger: more candidates and pairs
              scenes.mixed_fills(60, extent=(512, 384), size=(10.0, 60.0), seed=24, types=(0, 1))]
    rnd = cr.Renderer()
    rnd.resize_internal_buffers(512, 384)
    rnd.set_pipelining(True)
    batch = None
    for scene in frames:   # back to back: every rebuild alternates between the two sets of arrays
        batch = _render_scene(cr, rnd, scene, batch)
    ref_color, ref_stencil, ref_covered = _oracle_frame(oracle, rnd, frames[-1])
    assert np.array_equal(rnd.read_stencil(), ref_stencil) and np.array_equal(rnd.read_color().view(np.uint32), ref_color.view(np.uint32))
    assert int(rnd.stats().covered_samples) == ref_covered
    for scene in frames:   # one frame at a time
        batch = _render_scene(cr, rnd, scene, batch)
        ref_color, ref_stencil, _ = _oracle_frame(oracle, rnd, scene)
        assert np.array_equal(rnd.read_stencil(), ref_stencil) and np.array_equal(rnd.read_color().view(np.uint32), ref_color.view(np.uint32))
        ref = oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[3]), int(scene.shape_path_begin[4]))
        assert np.array_equal(batch[3].vertex_buffer(), ref.vertex_buffer) and np.array_equal(batch[3].index_buffer(), ref.index_buffer)
    batch.close()
    rnd.set_pipelining(False)
    batch = _render_scene(cr, rnd, frames[0])
    ref_color, ref_stencil, _ = _oracle_frame(oracle, rnd, frames[0])
    assert np.array_equal(rnd.read_stencil(), ref_stencil) and np.array_equal(rnd.read_color().view(np.uint32), ref_color.view(np.uint32))
    batch.close()
    rnd.close()


@pytest.mark.timeout(120)
def test_a_smaller_pass_after_resize_reuses_the_sort_scratch(cr, oracle):
    """One renderer, a large pass, then resize_internal_buffers and a much smaller pass (what __graft_entry__.smoke() does): the
    pair capacity SHRINKS, so the radix sort lays its histograms and the scan's ticket counters out differently inside the same
    scratch allocation. (A ticket counter found where histogram values used to be made the look-back scan wait for ever.)"""
    rnd = cr.Renderer()
    for scene in (scenes.mixed_fills(64, extent=(256, 192), rational=True), scenes.closed_cubic_strokes(24, extent=(256, 192)),
                  scenes.mixed_fills(300, extent=(512, 384), rational=True, seed=9), scenes.closed_cubic_strokes(6, extent=(128, 96))):
        rnd.resize_internal_buffers(scene.width, scene.height)
        batch = cr.ShapeBatch(rnd, scene.dynamic_stroke_options, scene.paths, scene.shape_path_begin)
        cmds = scenes.stencil_cover_commands(scene.n_shapes)
        rp = rnd.begin_render_pass()
        rp.set_instances(scene.transforms(), scene.colors)
        rp.render_batch(batch, cmds)
        rp.submit()
        color, stencil = rnd.read_color(), rnd.read_stencil()
        refs = [oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, int(scene.shape_path_begin[i]), int(scene.shape_path_begin[i + 1]))
                for i in range(scene.n_shapes)]
        ocmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in cmds]
        ref_color, ref_stencil, _, _ = oracle.render(rnd.config.to_c(), scene.width, scene.height, refs, ocmds, scene.transforms(), scene.colors, threads=4)
        assert np.array_equal(stencil, ref_stencil) and np.array_equal(color.view(np.uint32), ref_color.view(np.uint32))
        batch.close()
    rnd.close()


@pytest.mark.timeout(120)
def test_asynchronous_frame_readback(cr, oracle):
    """cr_renderer_read_color_texels_async: the frame of the pass submitted last arrives in caller memory while later passes run —
    also when that pass turns out to be undersized and is re-submitted (the snapshot taken behind the skipped attempt is replaced),
    with two read-backs in flight, and with the frame of an EARLIER pass waited for after a later pass has been submitted."""
    small = scenes.mixed_fills(12, extent=(512, 384), size=(10.0, 40.0), seed=5)
    large = scenes.mixed_fills(500, extent=(512, 384), size=(10.0, 120.0), rational=True, seed=6)
    rnd = cr.Renderer(cr.Configuration(color_format=cr.ColorFormat.Rgba8Unorm))
    rnd.resize_internal_buffers(512, 384)
    want = {}
    for name, scene in (("small", small), ("large", large)):
        b = _render_scene(cr, rnd, scene)
        frame = np.empty((384, 512), np.uint32)
        rnd.read_color_texels(frame.ctypes.data, frame.nbytes)
        want[name] = frame
        b.close()
    assert not np.array_equal(want["small"], want["large"])
    rnd.close()
    rnd = cr.Renderer(cr.Configuration(color_format=cr.ColorFormat.Rgba8Unorm))
    rnd.resize_internal_buffers(512, 384)
    got = [np.zeros((384, 512), np.uint32) for _ in range(4)]
    batches, tickets = [], []
    for k, scene in enumerate((small, large, small, large)):   # the large pass after a small one exceeds the capacities: re-submission
        batches.append(_render_scene(cr, rnd, scene))
        tickets.append(rnd.read_color_texels_async(got[k].ctypes.data, got[k].nbytes))
        if k >= 1:
            rnd.wait_readback(tickets[k - 1])                   # the frame before: its pass has been settled by this submit
            assert np.array_equal(got[k - 1], want["small" if (k - 1) % 2 == 0 else "large"]), f"frame {k - 1}"
    rnd.wait_readback(tickets[-1])
    assert np.array_equal(got[3], want["large"])
    for b in batches:
        b.close()
    rnd.close()
