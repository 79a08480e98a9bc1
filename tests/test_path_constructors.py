"""Host logic of the path constructors / conversions (src/path.rs:376-815) and the utils.rs helpers they use
(SURVEY section 8 rows f2, f4). The reference has no tests for them; these pin them by their mathematical contract:
curves are evaluated in float64 from the stored control points and compared with the geometry they are documented
to describe."""
import math

import numpy as np
import pytest

from contrast_renderer_b200 import utils as U
from contrast_renderer_b200.path import (IntegralCubicCurveSegment, IntegralQuadraticCurveSegment, Path, RationalCubicCurveSegment,
                                         RationalQuadraticCurveSegment, SegmentType)


def bernstein(n, i, t):
    return math.comb(n, i) * t ** i * (1 - t) ** (n - i)


def walk(path: Path):
    """Yields (kind, weights, control points incl. the segment's start) per segment in path order."""
    cursors = [0] * 5
    previous = np.asarray(path.start, np.float64)
    stores = path._stores()
    for kind in path.segment_types:
        seg = np.asarray(stores[kind][cursors[kind]], np.float64)
        cursors[kind] += 1
        nw = Path._N_WEIGHTS[kind]
        pts = np.vstack([previous, seg[nw:].reshape(-1, 2)])
        if kind == SegmentType.RationalQuadraticCurve:
            w = np.array([1.0, seg[0], 1.0])
        elif kind == SegmentType.RationalCubicCurve:
            w = seg[0:4]
        else:
            w = np.ones(len(pts))
        yield kind, w, pts
        previous = pts[-1]


def evaluate(w, pts, t):
    n = len(pts) - 1
    b = np.array([bernstein(n, i, t) for i in range(n + 1)]) * w
    return (b[:, None] * pts).sum(0) / b.sum()


def sample(path: Path, per_segment=9):
    out = []
    for _, w, pts in walk(path):
        out += [evaluate(w, pts, t) for t in np.linspace(0, 1, per_segment)]
    return np.array(out)


def mixed_path():
    p = Path([0.5, -1.0])
    p.push_line([2.0, 0.0])
    p.push_integral_quadratic_curve(IntegralQuadraticCurveSegment([[3.0, 1.5], [2.0, 3.0]]))
    p.push_rational_quadratic_curve(RationalQuadraticCurveSegment(0.6, [[0.5, 3.5], [-1.0, 2.0]]))
    p.push_integral_cubic_curve(IntegralCubicCurveSegment([[-2.0, 1.0], [-2.5, -0.5], [-1.0, -1.5]]))
    p.push_rational_cubic_curve(RationalCubicCurveSegment([1.0, 0.7, 1.8, 1.0], [[-0.5, -2.5], [0.2, -2.0], [0.4, -1.2]]))
    return p


# ------------------------------------------------------------------------------------------------ utils.rs
def test_motors_reproduce_standard_matrices():
    """translate2d / rotate2d through motor2d_to_mat3 (src/utils.rs:121-165)."""
    t = U.motor2d_to_mat3(U.translate2d([3.0, -4.0]))
    assert np.allclose(t, [[1, 0, 0], [0, 1, 0], [3, -4, 1]], atol=1e-6)
    a = 0.7
    r = U.motor2d_to_mat3(U.rotate2d(a))
    assert np.allclose(r, [[math.cos(a), math.sin(a), 0], [-math.sin(a), math.cos(a), 0], [0, 0, 1]], atol=1e-6)
    # composition: rotate first, then translate
    m = U.motor_product(U.translate2d([3.0, -4.0]), U.rotate2d(a))
    p = U.point_to_vec(U.motor_transform_point(m, U.vec_to_point([2.0, 1.0])))
    assert np.allclose(p, [3 + 2 * math.cos(a) - math.sin(a), -4 + 2 * math.sin(a) + math.cos(a)], atol=1e-5)
    assert abs(U.rotation2d(m) - a) < 1e-6
    assert np.allclose(U.translation2d(U.translate2d([3.0, -4.0])), [3.0, -4.0], atol=1e-6)
    # inverse and plane transformation: direction vectors rotate like points, a line stays incident with its points
    assert np.allclose(U.motor_product(m, U.motor_inverse(m)), [1, 0, 0, 0], atol=1e-6)
    line = np.array([-1.0, 0.5, 2.0], np.float32)   # 0.5 x + 2 y - 1 = 0 contains (2, 0)
    moved_line, moved_point = U.motor_transform_plane(m, line), U.motor_transform_point(m, U.vec_to_point([2.0, 0.0]))
    assert abs(float(np.dot(moved_line, moved_point))) < 1e-5


def test_line_line_intersection_and_colour_helpers():
    p = U.line_line_intersection([-1.0, 1.0, 0.0], [-2.0, 0.0, 1.0])   # x = 1 and y = 2
    assert np.allclose(U.point_to_vec(p), [1.0, 2.0], atol=1e-6)
    c = np.array([0.2, 0.5, 0.9, 0.7], np.float32)
    assert np.allclose(U.linear_to_srgb(U.srgb_to_linear(c)), c, atol=1e-6)
    assert abs(U.srgb_to_linear([0.5, 0, 0, 1])[0] - 0.21404114) < 1e-6 and U.srgb_to_linear(c)[3] == c[3]
    persp = U.perspective_projection(math.pi / 2, 2.0, 1.0, 10.0)
    assert np.allclose(persp[0][0], 0.5) and np.allclose(persp[1][1], 1.0) and persp[2][3] == 1.0
    ident = np.eye(4, dtype=np.float32)
    assert np.allclose(U.matrix_multiplication(persp, ident), persp) and np.allclose(U.matrix_multiplication(ident, persp), persp)
    assert np.allclose(U.complex_powi([0.6, 0.8], 5), U.complex_powf([0.6, 0.8], 5.0), atol=1e-5)


# ------------------------------------------------------------------------------------------------ path.rs:387-617
def test_transform_moves_every_control_point():
    p, q = mixed_path(), mixed_path()
    motor = U.motor_product(U.translate2d([1.5, -2.0]), U.rotate2d(0.9))
    q.transform(1.0, motor)
    expect = np.array([U.point_to_vec(U.motor_transform_point(motor, U.vec_to_point(v))) for v in sample(p)])
    assert np.allclose(sample(q), expect, atol=2e-5)
    s = mixed_path()
    s.transform(2.5, U.translate2d([1.0, 1.0]))   # rotation-free motor: uniform scale, then translation
    assert np.allclose(sample(s), sample(p) * 2.5 + 1.0, atol=2e-5)
    assert s.segment_types == p.segment_types and np.allclose(s.rational_cubic_curve_segments[0][:4], [1.0, 0.7, 1.8, 1.0])


def test_reverse_is_an_involution_that_swaps_the_ends():
    p, q = mixed_path(), mixed_path()
    q.reverse()
    assert np.array_equal(q.start, p.get_end()) and np.array_equal(q.get_end(), p.start)
    assert q.segment_types == p.segment_types[::-1]
    assert np.allclose(sample(q), sample(p)[::-1], atol=1e-6)
    q.reverse()
    assert np.array_equal(q.start, p.start) and q.segment_types == p.segment_types
    for a, b in zip(p._stores(), q._stores()):
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


def test_conversions_keep_the_curve():
    p, q = mixed_path(), mixed_path()
    q.convert_integral_curves_to_rational_curves()
    assert not q.integral_quadratic_curve_segments and not q.integral_cubic_curve_segments
    assert q.segment_types == [0, 3, 3, 4, 4] and len(q.rational_quadratic_curve_segments) == 2 and len(q.rational_cubic_curve_segments) == 2
    assert np.array_equal(q.rational_quadratic_curve_segments[0], np.float32([1.0, 3.0, 1.5, 2.0, 3.0]))   # path order is kept
    assert np.allclose(sample(q), sample(p), atol=1e-6)
    r = mixed_path()
    r.convert_quadratic_curves_to_cubic_curves()
    assert not r.integral_quadratic_curve_segments and not r.rational_quadratic_curve_segments
    assert r.segment_types == [0, 2, 4, 2, 4]
    assert np.allclose(sample(r), sample(p), atol=2e-6)
    assert np.allclose(r.rational_cubic_curve_segments[0][[0, 3]], 1.0)
    both = mixed_path()
    both.convert_quadratic_curves_to_cubic_curves()
    both.convert_integral_curves_to_rational_curves()
    assert both.segment_types == [0, 4, 4, 4, 4] and np.allclose(sample(both), sample(p), atol=2e-6)


def test_append_and_close():
    a, b = mixed_path(), Path([9.0, 9.0])
    b.push_line([5.0, 5.0])
    n = len(a.segment_types)
    a.append(b)
    assert len(a.segment_types) == n + 1 and not b.segment_types and not b.line_segments and np.array_equal(a.get_end(), [5.0, 5.0])
    a.close()
    assert np.array_equal(a.get_end(), a.start)
    m = len(a.segment_types)
    a.close()   # already closed: no-op (src/path.rs:621-623)
    assert len(a.segment_types) == m


# ------------------------------------------------------------------------------------------------ path.rs:629-815
@pytest.mark.parametrize("large_arc", [False, True])
@pytest.mark.parametrize("sweep", [False, True])
@pytest.mark.parametrize("rotation", [0.0, 0.6])
def test_elliptical_arc_follows_the_svg_rules(large_arc, sweep, rotation):
    """src/path.rs:638-703 against the SVG implementation notes it cites: every emitted rational quadratic lies on one
    ellipse with the given radii and rotation through both end points; large_arc picks the > 180 degree arc; sweep is
    SVG's "positive angle" on a y-down screen: the reference takes the chord as to - from where the notes use
    (from - to) / 2 (src/path.rs:648), which mirrors the centre, so in its own y-up model frame sweep = true turns
    clockwise (and looks like SVG once the frame is shown y-down)."""
    start, to, radii = np.array([1.0, 0.5]), np.array([3.0, 2.0]), np.array([2.5, 1.5])
    p = Path(start)
    p.push_elliptical_arc(radii, rotation, large_arc, sweep, to)
    assert p.segment_types and all(k == SegmentType.RationalQuadraticCurve for k in p.segment_types)
    assert len(p.segment_types) <= 3 and np.allclose(p.get_end(), to, atol=1e-5)
    pts = sample(p, 17)
    # recover the centre: all samples satisfy |R^-1 (q - c) / radii| = 1; solve the linear system of the conic through them
    c, s = math.cos(rotation), math.sin(rotation)
    local = np.stack([(pts[:, 0] * c + pts[:, 1] * s) / radii[0], (-pts[:, 0] * s + pts[:, 1] * c) / radii[1]], 1)
    a = np.hstack([2 * local, np.ones((len(local), 1))])
    sol, *_ = np.linalg.lstsq(a, (local ** 2).sum(1), rcond=None)
    centre = sol[:2]
    radius = np.linalg.norm(local - centre, axis=1)
    assert np.allclose(radius, 1.0, atol=2e-5), "samples are not on one ellipse with the requested radii"
    ang = np.unwrap(np.arctan2(local[:, 1] - centre[1], local[:, 0] - centre[0]))
    steps = np.diff(ang)
    assert (steps > -1e-6).all() or (steps < 1e-6).all(), "arc does not turn monotonically"
    swept = ang[-1] - ang[0]
    assert (abs(swept) > math.pi) == large_arc
    assert (swept < 0) == sweep
    # each piece spans at most 120 degrees (src/path.rs:676-677)
    assert abs(swept) / len(p.segment_types) <= 2 * math.pi / 3 + 1e-5


def test_elliptical_arc_degenerate_and_undersized_radii():
    p = Path([0.0, 0.0])
    p.push_elliptical_arc([0.0, 1.0], 0.0, False, True, [2.0, 0.0])   # zero radius: a line (src/path.rs:641-644)
    assert p.segment_types == [SegmentType.Line] and np.array_equal(p.get_end(), [2.0, 0.0])
    q = Path([0.0, 0.0])
    q.push_elliptical_arc([0.5, 0.5], 0.0, False, True, [4.0, 0.0])   # radii scaled up to the half distance: a half circle
    pts = sample(q, 9)
    assert np.allclose(np.linalg.norm(pts - [2.0, 0.0], axis=1), 2.0, atol=1e-5) and np.allclose(q.get_end(), [4.0, 0.0], atol=1e-5)


def test_shape_constructors():
    rect = Path.from_rect([1.0, 2.0], [3.0, 0.5])
    assert np.array_equal(rect.start, [-2.0, 1.5]) and rect.segment_types == [0, 0, 0]
    assert np.array_equal(np.array(rect.line_segments), [[-2.0, 2.5], [4.0, 2.5], [4.0, 1.5]])
    poly = Path.from_regular_polygon([1.0, 1.0], 2.0, 0.25, 7)
    v = np.vstack([poly.start, np.array(poly.line_segments)])
    assert len(v) == 7 and np.allclose(np.linalg.norm(v - 1.0, axis=1), 2.0, atol=1e-6)
    assert np.allclose(np.arctan2(v[0, 1] - 1, v[0, 0] - 1), 0.25, atol=1e-6)
    circle = Path.from_circle([1.0, -1.0], 2.5)
    assert circle.segment_types == [3, 3, 3, 3] and np.array_equal(circle.start, circle.get_end())
    assert np.allclose(np.linalg.norm(sample(circle, 33) - [1.0, -1.0], axis=1), 2.5, atol=1e-6)
    ell = sample(Path.from_ellipse([0.0, 0.0], [3.0, 1.0]), 33)
    assert np.allclose((ell[:, 0] / 3.0) ** 2 + ell[:, 1] ** 2, 1.0, atol=1e-6)
    rr = Path.from_rounded_rect([0.0, 0.0], [2.0, 1.0], 0.25)
    assert rr.segment_types == [0, 3] * 4 and np.array_equal(rr.start, rr.get_end())
    pts = sample(rr, 17)
    inner = np.maximum(np.abs(pts) - [1.75, 0.75], 0.0)   # distance to the inner rectangle is the corner radius on the arcs
    on_arc = (inner > 1e-6).all(1)
    assert on_arc.any() and np.allclose(np.linalg.norm(inner[on_arc], axis=1), 0.25, atol=1e-6)
    assert (np.abs(pts[:, 0]) <= 2.0 + 1e-6).all() and (np.abs(pts[:, 1]) <= 1.0 + 1e-6).all()
    # the reference's constructors run clockwise in a y-up frame (src/path.rs:731-738): negative shoelace area
    x, y = v[:, 0], v[:, 1]
    assert 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y)) > 0   # regular polygon: counter-clockwise (increasing angle)
    rv = np.vstack([rect.start, np.array(rect.line_segments)])
    assert 0.5 * float(np.sum(rv[:, 0] * np.roll(rv[:, 1], -1) - np.roll(rv[:, 0], -1) * rv[:, 1])) < 0


# ---------------------------------------------------------------------------------------------- demo helpers (f4)
def test_motor3d_embedding_agrees_with_the_2d_motor():
    """src/utils.rs:149-180: `motor3d_to_mat4(motor2d_to_motor3d(m))` moves (x, y, 0, 1) like `motor2d_to_mat3(m)` moves (x, y, 1)."""
    rng = np.random.default_rng(7)
    for _ in range(8):
        m2 = U.motor_product(U.translate2d(rng.uniform(-2, 2, 2)), U.rotate2d(float(rng.uniform(-3, 3))))
        m4, m3 = U.motor3d_to_mat4(U.motor2d_to_motor3d(m2)), U.motor2d_to_mat3(m2)
        x, y = rng.uniform(-1, 1, 2).astype(np.float32)
        p3 = m4[3] + x * m4[0] + y * m4[1]
        p2 = m3[2] + x * m3[0] + y * m3[1]
        assert np.allclose(p3[[0, 1, 3]], p2, atol=1e-5) and abs(p3[2]) < 1e-6


def test_camera_motor_of_the_showcase():
    """examples/showcase/main.rs:163-201: Translator(1, 0, 0, -d / 2) * view_rotation pushes the scene d along +z (the projection's
    clip w is the view-space z, src/utils.rs:183-192), instance motors (.., tx, ty, tz) translate by -2 (tx, ty, tz)."""
    view = U.motor3d_product(U.translator3d(0.0, 0.0, -0.5 * 10.0), U.rotor3d_to_motor3d(U.rotate_around_axis(0.0, [0, 1, 0])))
    m = U.motor3d_to_mat4(view)
    assert np.allclose(m[3], [0, 0, 10, 1], atol=1e-6) and np.allclose(m[:3, :3], np.eye(3), atol=1e-6)
    # a quarter turn about z takes the x axis to the y axis (same sense as rotate2d)
    q = U.motor3d_to_mat4(U.rotor3d_to_motor3d(U.rotate_around_axis(np.pi / 2, [0, 0, 1])))
    assert np.allclose(q[0], [0, 1, 0, 0], atol=1e-6) and np.allclose(q[1], [-1, 0, 0, 0], atol=1e-6)
    # rotations about x and y are proper rotations of the same handedness (cyclic continuation)
    for axis, src, dst in (([1, 0, 0], 1, 2), ([0, 1, 0], 2, 0)):
        r = U.motor3d_to_mat4(U.rotor3d_to_motor3d(U.rotate_around_axis(np.pi / 2, axis)))
        expect = np.zeros(4, np.float32)
        expect[dst] = 1.0
        assert np.allclose(r[src], expect, atol=1e-6) and np.isclose(np.linalg.det(r[:3, :3].astype(np.float64)), 1.0, atol=1e-5)
    # the projected origin of the instance grid's centre cell sits in front of the camera
    proj = U.matrix_multiplication(U.perspective_projection(np.pi * 0.5, 16 / 9, 1.0, 1000.0), m)
    clip = proj[3]
    assert clip[3] > 0 and abs(clip[0]) < 1e-6 and abs(clip[1]) < 1e-6


def test_save_png_round_trip(tmp_path):
    import struct
    import zlib
    frame = np.zeros((5, 7, 4), np.float32)
    frame[1:4, 2:5] = [0.5, 0.25, 0.0, 0.5]   # premultiplied: colour (1, 0.5, 0) at half opacity
    path = tmp_path / "frame.png"
    U.save_png(str(path), frame)
    data = path.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    w, h, depth, kind = struct.unpack(">IIBB", data[16:26])
    assert (w, h, depth, kind) == (7, 5, 8, 6)
    idat_len = struct.unpack(">I", data[33:37])[0]
    raw = zlib.decompress(data[41:41 + idat_len])
    rows = np.frombuffer(raw, np.uint8).reshape(5, 1 + 7 * 4)
    assert (rows[:, 0] == 0).all()
    px = rows[:, 1:].reshape(5, 7, 4)
    assert tuple(px[2, 3]) == (255, 188, 0, 128) and tuple(px[0, 0]) == (0, 0, 0, 0)
    texels = np.arange(5 * 7 * 4, dtype=np.uint8).reshape(5, 7, 4)
    U.save_png(str(path), texels)
    data = path.read_bytes()
    idat_len = struct.unpack(">I", data[33:37])[0]
    assert np.array_equal(np.frombuffer(zlib.decompress(data[41:41 + idat_len]), np.uint8).reshape(5, 29)[:, 1:].reshape(5, 7, 4), texels)
