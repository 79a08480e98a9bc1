"""Decision test for reference quirk C.9: non-zero fills of cubic curves are exact for one
sense of traversal and overfill the chord / curve sliver for the other.

"The oracle and the kernels agree" cannot decide whether that is the reference's behaviour or a sign error shared by both
(the geometric products live in the un-vendored crate geometric_algebra 0.3.0). This file decides it from statements the
reference makes about ITSELF, none of which needs the crate's source:

 (A) sign of (A v B) v C. src/utils.rs:80-101 `do_convex_polygons_overlap` "expects the vertices to be ordered clockwise",
     is fed `aabb_to_convex_polygon` = (x0,y0),(x0,y1),(x1,y1),(x1,y0) (clockwise with y up), and treats
     `point v (a[i+1] v a[i]) > 0` for EVERY point as "edge i separates". So for a clockwise polygon a point OUTSIDE edge i
     has a positive product: (A v B) v C = -det[A; B; C].
 (B) normal stored by a tangent Plane. src/stroke.rs:272-281 builds the start cap of an open stroke at
     `offset_control_point(p, rotate_90_degree_clockwise(tangent), width / 2)`, and src/utils.rs:104-106 spells
     rotate_90_degree_clockwise out in components: (g1, g2) -> (g2, -g1). A start cap lies BEHIND the start point, so
     (g1, g2) of the tangent P v Q is the direction of travel turned 90 degrees clockwise: the right-hand normal.

From (A): `emit_cubic_curve_triangle!` reverses a triangle when its product is negative (src/fill.rs:127-129), so every
Loop-Blinn triangle ends up with a positive product = CLOCKWISE (y up) whatever the sense of the path; under an
orientation-preserving instance matrix that is a back face (front = Ccw, src/renderer.rs:477) and the fill stencil state
DECREMENTS on it (src/renderer.rs:577-582).
From (B): `normalize_implicit_curve_side` (src/fill.rs:98-114) flips k, l until tangent . gradient <= 0, i.e. until
k^3 - lmn decreases towards the tangent's normal; the fragment test keeps k^3 - lmn <= 0 (src/shaders.wgsl:245-249): the
kept side is always the RIGHT of the direction of travel, and a control point joins the solid fan exactly when it lies on
the kept side (implicit value < 0, src/fill.rs:193-198).

Worked by hand for the arch P0 (0,0), P1 (1,1.2), P2 (2,1.2), P3 (3,0) closed over the top by (3,2.5), (0,2.5), y up:
 * sense A, counter-clockwise (travel +x along the arch, interior on the left): kept side = below the arch = the sliver
   between chord and curve, OUTSIDE the shape; control points are on the other side, so the fan is the rectangle, drawn as a
   strip whose triangles are clockwise (src/vertex.rs:28-35 turns the fan around) => -1 inside the rectangle; the sliver gets
   another -1 from the clockwise curve triangles => -2: even-odd is right, non-zero OVERFILLS the sliver.
 * sense B, clockwise (travel -x along the arch, interior on the right): kept side = above the arch, between the curve and
   its control polygon, INSIDE the shape; the control points join the fan, whose polygon therefore stops at the control
   polygon => +1 there, -1 between control polygon and curve, 0 in the sliver: non-zero coverage is EXACT.
The numbers below are computed from that derivation with numpy alone and compared with the oracle's stencil buffer.
"""
import numpy as np

from contrast_renderer_b200.path import Path, PathSoA
from contrast_renderer_b200.renderer import Configuration

P0, P1, P2, P3 = (0.0, 0.0), (1.0, 1.2), (2.0, 1.2), (3.0, 0.0)
TOP_RIGHT, TOP_LEFT = (3.0, 2.5), (0.0, 2.5)
W, H = 320, 280
X0, X1, Y0, Y1 = -0.5, 3.5, -0.5, 3.0   # model window mapped to the target, y UP (orientation preserving)


def test_ga_sign_conventions_are_pinned_by_the_reference_itself(oracle):
    # (A) src/utils.rs:70-101 on the clockwise AABB polygon
    x0, y0, x1, y1 = 1.0, 2.0, 4.0, 3.0
    poly = [(x0, y0), (x0, y1), (x1, y1), (x1, y0)]
    outside = [(0.0, 2.5), (2.0, 4.0), (5.0, 2.5), (2.0, 1.0)]   # beyond edge 0 (left), 1 (top), 2 (right), 3 (bottom)
    inside = (2.0, 2.5)
    for i in range(4):
        a, b = poly[(i + 1) % 4], poly[i]
        assert oracle.ga_triple(a, b, outside[i]) > 0.0    # `point v plane <= 0 => continue` must NOT trigger for a separated point
        assert oracle.ga_triple(a, b, inside) < 0.0
    # (B) src/stroke.rs:272-281 + src/utils.rs:104-106: rotate_90_degree_clockwise(tangent) points against the direction of travel
    for p, q in [((0.0, 0.0), (2.0, 0.0)), ((1.0, 1.0), (0.0, 3.0)), ((2.0, -1.0), (-1.0, -2.0))]:
        g = oracle.ga_join(p, q)
        rotated = np.array([g[2], -g[1]])
        travel = np.array(q) - np.array(p)
        assert np.dot(rotated, travel) < 0.0 and abs(rotated[0] * travel[1] - rotated[1] * travel[0]) < 1e-6
        assert travel[0] * g[2] - travel[1] * g[1] < 0.0              # (g1, g2) is to the right of the direction of travel


def _transform():
    m = np.zeros(16, np.float32)   # column vectors: clip = col0 x + col1 y + col3 (src/shaders.wgsl:20-27,72); no mirroring
    m[0], m[5], m[10], m[15] = 2.0 / (X1 - X0), 2.0 / (Y1 - Y0), 1.0, 1.0
    m[12], m[13] = -1.0 - X0 * m[0], -1.0 - Y0 * m[5]
    return m


def _sample_positions():
    ys, xs = np.mgrid[0:H, 0:W]
    ndc_x, ndc_y = (xs + 0.5) / W * 2.0 - 1.0, 1.0 - (ys + 0.5) / H * 2.0
    return X0 + (ndc_x + 1.0) * 0.5 * (X1 - X0), Y0 + (ndc_y + 1.0) * 0.5 * (Y1 - Y0)


def _inside(poly, x, y):
    """Even-odd point-in-polygon (float64), vectorised over the sample grid."""
    inside = np.zeros(x.shape, bool)
    pts = np.asarray(poly, np.float64)
    for (ax, ay), (bx, by) in zip(pts, np.roll(pts, -1, axis=0)):
        crosses = (ay > y) != (by > y)
        with np.errstate(divide="ignore", invalid="ignore"):
            xi = ax + (y - ay) * (bx - ax) / (by - ay)
        inside ^= crosses & (x < xi)
    return inside


def _distance(poly, x, y, closed=True):
    pts = np.asarray(poly, np.float64)
    nxt = np.roll(pts, -1, axis=0)
    if not closed:
        pts, nxt = pts[:-1], nxt[:-1]
    best = np.full(x.shape, np.inf)
    for (ax, ay), (bx, by) in zip(pts, nxt):
        dx, dy = bx - ax, by - ay
        t = np.clip(((x - ax) * dx + (y - ay) * dy) / (dx * dx + dy * dy), 0.0, 1.0)
        best = np.minimum(best, np.hypot(x - (ax + t * dx), y - (ay + t * dy)))
    return best


def _render(oracle, path, winding_bits):
    soa = PathSoA.from_paths([path])
    shape = oracle.shape_from_paths([], soa)
    cfg = Configuration(winding_counter_bits=winding_bits, clip_nesting_counter_bits=4).to_c()
    _, stencil, _, _ = oracle.render(cfg, W, H, [shape], [(0, 0, 1, 0, 0, 0, 0)], _transform().reshape(1, 16), None)
    return stencil[..., 0].astype(np.int64)


def test_cubic_fill_stencil_follows_the_hand_derivation(oracle):
    x, y = _sample_positions()
    t = np.linspace(0.0, 1.0, 513)[:, None]
    arch = ((1 - t) ** 3 * np.array(P0) + 3 * t * (1 - t) ** 2 * np.array(P1) + 3 * t ** 2 * (1 - t) * np.array(P2) + t ** 3 * np.array(P3))
    rectangle = [P0, P3, TOP_RIGHT, TOP_LEFT]
    sliver = _inside(list(arch), x, y)                                           # between chord and curve (outside the shape)
    under_control_polygon = _inside([P0, P1, P2, P3], x, y)
    between = under_control_polygon & ~sliver                                    # between the curve and its control polygon (inside the shape)
    rect = _inside(rectangle, x, y)
    fan_b = rect & ~under_control_polygon                                        # sense B's fan stops at the control polygon
    margin = 2.5 * (X1 - X0) / W                                                 # stay clear of every edge any triangle can have
    far = np.ones(x.shape, bool)
    for poly in (rectangle, [P0, P1, P2, P3], [P0, P1, P3], [P0, P2, P3], [P1, P2, P3], [P0, P1, P2], [P0, P2], [P1, P3]):
        far &= _distance(poly, x, y) > margin
    far &= _distance(list(arch), x, y, closed=False) > margin
    assert far[sliver].sum() > 1000 and far[between].sum() > 500 and far[fan_b].sum() > 5000

    a = Path(P0)   # sense A: counter-clockwise, y up
    a.push_integral_cubic_curve([P1, P2, P3])
    a.push_line(TOP_RIGHT)
    a.push_line(TOP_LEFT)
    a.push_line(P0)
    b = Path(P0)   # sense B: the same outline the other way round
    b.push_line(TOP_LEFT)
    b.push_line(TOP_RIGHT)
    b.push_line(P3)
    b.push_integral_cubic_curve([P2, P1, P0])

    want_a = np.where(rect, -1, 0) + np.where(sliver, -1, 0)
    want_b = np.where(fan_b, 1, 0) + np.where(between, -1, 0)
    got_a, got_b = _render(oracle, a, 4), _render(oracle, b, 4)
    assert np.array_equal((got_a % 16)[far], (want_a % 16)[far])
    assert np.array_equal((got_b % 16)[far], (want_b % 16)[far])
    shape = rect & ~sliver                                                       # the region the outline encloses
    assert np.array_equal((got_b != 0)[far], shape[far])                         # clockwise: non-zero coverage exact
    assert np.array_equal((got_a != 0)[far], rect[far]) and (got_a[far & sliver] % 16 == 14).all()   # counter-clockwise: the sliver is overfilled
    # even-odd (one winding bit) is exact for both senses
    for path in (a, b):
        assert np.array_equal((_render(oracle, path, 1) & 1 != 0)[far], shape[far])
