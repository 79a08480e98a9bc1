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
