"""Pins the CPU oracle. The reference has no tests, golden vectors or fixtures and cannot be built here (Rust), so the
oracle's parity is UNPINNED by the reference itself; these are the hand-derivable anchors of SURVEY.md §8c and §4 plus
independent numerical cross-checks (numpy float64) of the arithmetic that lives in the reference's un-vendored
dependencies (geometric_algebra 0.3.0 closed forms, polynomial solvers, libm)."""
import math

import numpy as np
import pytest

from contrast_renderer_b200 import _abi, scenes
from contrast_renderer_b200.path import (Cap, CurveApproximation, DashInterval, DynamicStrokeOptions, Join, Path, PathSoA, RationalQuadraticCurveSegment,
                                         StrokeOptions)
from contrast_renderer_b200.renderer import Configuration, orthographic_transform

SOLID = [DynamicStrokeOptions.Solid(Join.Miter, Cap.Butt, Cap.Butt)]


# ------------------------------------------------------------------------------------------ arithmetic contract
def test_elementary_functions_are_correctly_rounded(oracle):
    """cr_arith.h replaces libm so that g++ and nvcc agree bit for bit; it must still BE atan2/acos/sincos/pow."""
    rng = np.random.default_rng(1)
    lib = oracle.lib()
    worst = 0.0
    for _ in range(4000):
        y, x = np.float32(rng.normal()), np.float32(rng.normal())
        got = lib.oracle_atan2(y, x)
        want = math.atan2(float(y), float(x))
        worst = max(worst, abs(got - want) / np.spacing(np.float32(abs(want)) + np.float32(1e-30)))
        c = np.float32(rng.uniform(-1, 1))
        worst = max(worst, abs(lib.oracle_acos(c) - math.acos(float(c))) / np.spacing(np.float32(math.acos(float(c)))))
        b, e = np.float32(rng.uniform(0.01, 4)), np.float32(rng.uniform(-3, 3))
        want = float(b) ** float(e)
        worst = max(worst, abs(lib.oracle_pow(b, e) - want) / np.spacing(np.float32(want)))
    assert worst <= 0.51, worst   # half an ulp: rounded once from a binary64 evaluation
    assert lib.oracle_atan2(np.float32(0.0), np.float32(-1.0)) == np.float32(math.pi)
    assert lib.oracle_acos(np.float32(0.0)) == np.float32(math.pi / 2)
    assert lib.oracle_wgsl_mod(np.float32(-7.5), np.float32(2.0)) == np.float32(-1.5)   # WGSL % truncates (src/shaders.wgsl:211)


@pytest.mark.parametrize("degree", [1, 2, 3, 4])
def test_polynomial_solvers_find_the_roots(oracle, degree):
    """SURVEY Appendix B contract: ascending coefficients, roots as numerator / denominator, complex roots included."""
    rng = np.random.default_rng(degree)
    for _ in range(300):
        c = rng.uniform(-2, 2, degree + 1).astype(np.float32)
        if abs(c[-1]) < 0.2:
            continue
        disc, roots, _ = oracle.solve(c)
        want = np.roots(c[::-1].astype(np.float64))
        got = (roots[:, 0] + 1j * roots[:, 1]) / roots[:, 2]
        assert len(got) == degree
        for r in got:   # every returned root is a root (f32 evaluation noise scales with the conditioning)
            assert np.min(np.abs(want - r)) < 2e-2 * max(1.0, abs(r)), (c, got, want)
        if degree == 2:
            assert (disc > 0) == (abs(want[0].imag) < 1e-12 and abs(want[0] - want[1]) > 0)   # src/curve.rs:214


def test_cubic_discriminant_sign_is_loop_blinn(oracle):
    """> 0 three real roots (serpentine), < 0 one real root (loop): src/fill.rs:53-65."""
    disc, _, _ = oracle.solve([-6.0, 11.0, -6.0, 1.0])   # (x-1)(x-2)(x-3)
    assert disc > 0
    disc, roots, real = oracle.solve([1.0, 0.0, 0.0, 1.0])   # x^3 + 1: one real root -1
    assert disc < 0 and abs(roots[real, 0] / roots[real, 2] + 1.0) < 1e-5 and abs(roots[real, 1]) < 1e-5


# -------------------------------------------------------------------------------------------------- curve.rs
def test_quarter_circle_uniform_tangent_angle(oracle):
    """quarter circle (1,0)->(0,1), control (1,1), weight 1/sqrt(2), angle step 0.1: 16 steps, 15 strictly
    increasing symmetric parameters + the end, every sample on the unit circle, equal tangent-angle increments."""
    w = 1.0 / math.sqrt(2.0)
    cp = [[1, 0], [1, 1], [0, 1]]
    params = oracle.uniform_tangent_angle(_abi.CR_SEG_RATIONAL_QUADRATIC, cp, [w], 0.1)
    assert len(params) == 16 and params[-1] == 1.0
    inner = params[:-1]
    assert np.all(np.diff(inner) > 0) and 0 < inner[0] and inner[-1] < 1
    assert np.allclose(inner + inner[::-1], 1.0, atol=2e-6) and abs(inner[7] - 0.5) < 1e-6
    angles = []
    for t in params:
        xy, normal = oracle.curve_eval(_abi.CR_SEG_RATIONAL_QUADRATIC, cp, [w], float(t))
        assert abs(np.hypot(*xy) - 1.0) < 2e-6
        angles.append(math.atan2(normal[1], normal[0]))
    steps = np.diff(np.unwrap([math.atan2(0.0, 1.0)] + angles))   # start normal: direction (0,1) rotated clockwise = (1,0)
    assert np.allclose(np.abs(steps), (math.pi / 2) / 16, atol=1e-5)


@pytest.mark.parametrize("kind", [_abi.CR_SEG_INTEGRAL_QUADRATIC, _abi.CR_SEG_INTEGRAL_CUBIC, _abi.CR_SEG_RATIONAL_QUADRATIC, _abi.CR_SEG_RATIONAL_CUBIC])
def test_uniform_tangent_angle_invariants(oracle, kind):
    """§4 invariant 4: parameters ascending inside [0,1], last = 1, consecutive tangents differ by about angle_step: the step
    count is `(angle / angle_step + 0.5) as usize` (src/curve.rs:233), i.e. rounded to nearest, so an interval of k steps has
    increments of at most angle_step * (k + 0.5) / k <= 1.25 angle_step."""
    rng = np.random.default_rng(kind)
    step = 0.15
    checked = 0
    for _ in range(60):
        n = 3 if kind in (_abi.CR_SEG_INTEGRAL_QUADRATIC, _abi.CR_SEG_RATIONAL_QUADRATIC) else 4
        base = np.array([[0, 0], [1, 1.2], [2.2, 1.0], [3, -0.2]])[:n] if n == 4 else np.array([[0, 0], [1.2, 1.5], [2.5, 0]])
        cp = base + rng.uniform(-0.3, 0.3, (n, 2))
        wts = None
        if kind == _abi.CR_SEG_RATIONAL_QUADRATIC:
            wts = [rng.uniform(0.5, 2.0)]
        elif kind == _abi.CR_SEG_RATIONAL_CUBIC:
            wts = rng.uniform(0.6, 1.6, 4)
        params = oracle.uniform_tangent_angle(kind, cp, wts, step)
        assert params[-1] == 1.0 and np.all(params >= 0) and np.all(params <= 1)
        assert np.all(np.diff(params) >= -1e-6)
        prev = oracle.curve_eval(kind, cp, wts, 0.0)[1]
        for t in params:
            cur = oracle.curve_eval(kind, cp, wts, float(t))[1]
            turn = abs(math.atan2(prev[0] * cur[1] - prev[1] * cur[0], prev[0] * cur[0] + prev[1] * cur[1]))
            assert turn <= step * 1.25 + 2e-3, (kind, cp, params)
            prev = cur
            checked += 1
    assert checked > 300


# ------------------------------------------------------------------------------------------------- stroke.rs
def test_square_corner_miter_join(oracle):
    """SURVEY A.3 worked example: (0,0)->(1,0)->(1,1), width w, offset 0: side_sign = -1 for the left turn, edge vertices
    (1,-w/2) and (1+w/2,0), miter tip (1+w/2,-w/2), arc length grows by acos(0)/(2 pi) w = w/4."""
    w = 0.25
    p = Path([0, 0], StrokeOptions(width=w, offset=0.0, miter_clip=4.0, closed=False, curve_approximation=CurveApproximation.UniformTangentAngle(0.1)))
    p.push_line([1, 0])
    p.push_line([1, 1])
    shape = oracle.shape_from_paths(SOLID, PathSoA.from_paths([p]))
    line, joint = shape.vertices()[0], shape.vertices()[1]
    assert len(joint) == 5 and shape.indices()[1].tolist() == [0, 1, 2, 3, 4, 0xFFFF]
    assert np.allclose(joint["pos"], [[1, 0], [1, -w / 2], [1 + w / 2, 0], [1 + w / 2, -w / 2], [1 + w / 2, -w / 2]], atol=1e-6)
    assert np.allclose(joint["tex"][:, 2], 1.0 / w)            # offset along the path = length / width
    assert set(joint["flags"].tolist()) == {0}                    # joints never carry the end-cap flag (quirk C.2)
    # line strip: start cap tip pair, (no pair at the start of a first LINE segment, stroke.rs:285), corner pairs, end cap
    assert len(line) == 2 + 2 + 2 + 2 + 2 + 2
    assert np.allclose(line["pos"][0], [-w / 2, w / 2]) and np.allclose(line["pos"][1], [-w / 2, -w / 2])   # cap tip moved backwards by |w|/2
    assert np.allclose(line["tex"][0], [-0.5, -0.5]) and np.allclose(line["tex"][1], [0.5, -0.5])
    after_join = line[4]
    assert abs(after_join["tex"][1] - (1.0 + w / 4) / w) < 1e-5
    assert (line["flags"][-4:] == 0x10000).all() and (line["flags"][:-4] == 0).all()                       # end cap strip (stroke.rs:443-462)
    assert abs(line["tex"][-1][1] - (2.0 + w / 4 + w / 2) / w) < 1e-5
    # every pair is `width` apart and the hull contains every emitted vertex (§4 invariants 2, 3)
    pairs = line["pos"].reshape(-1, 2, 2)
    assert np.allclose(np.linalg.norm(pairs[:, 0] - pairs[:, 1], axis=1), w, atol=1e-6)
    hull = shape.vertices()[7]["pos"]
    assert hull.min(0)[0] <= line["pos"].min(0)[0] + 1e-6 and hull.max(0)[1] >= line["pos"].max(0)[1] - 1e-6


def test_stroke_strip_geometry(oracle):
    """§4 invariant 3 on curves: each vertex pair straddles the curve, |right - left| = width, arc length non-decreasing."""
    scene = scenes.closed_cubic_strokes(40)
    for i in range(scene.n_shapes):
        shape = oracle.shape_from_paths(scene.dynamic_stroke_options, scene.paths, i, i + 1)
        line = shape.vertices()[0]
        width = float(scene.paths.stroke_options["width"][i])
        pairs = line["pos"].reshape(-1, 2, 2)
        assert np.allclose(np.linalg.norm(pairs[:, 0] - pairs[:, 1], axis=1), width, rtol=2e-4, atol=2e-6)
        assert np.all(line["tex"][0::2, 0] == -0.5) and np.all(line["tex"][1::2, 0] == 0.5)
        along = line["tex"][0::2, 1]
        assert np.all(np.diff(along) >= -1e-5)
        idx = shape.indices()[0]
        assert idx[-1] == 0xFFFF and np.count_nonzero(idx == 0xFFFF) >= 1
        assert len(idx) == len(line) + np.count_nonzero(idx == 0xFFFF)


# --------------------------------------------------------------------------------------------------- fill.rs
def test_fixed_curve_weights(oracle):
    """integral quadratic weights (1,1),(1/2,0),(0,0) (src/fill.rs:288-292); rational quadratic
    (1,1,1),(1/(2w),0,1/w),(0,0,1) (src/fill.rs:324-328), whose implicit u^2 - v w vanishes on the exact circle."""
    w = np.float32(1.0 / math.sqrt(2.0))
    p = Path([1, 0])
    p.push_integral_quadratic_curve([[1, 1], [0, 1]])
    p.push_rational_quadratic_curve(RationalQuadraticCurveSegment(w, [[-1, 1], [-1, 0]]))
    p.push_line([1, 0])
    shape = oracle.shape_from_paths([], PathSoA.from_paths([p]))
    v = shape.vertices()
    assert v[3]["pos"].tolist() == [[0, 1], [1, 1], [1, 0]] and v[3]["w"].tolist() == [[1, 1], [0.5, 0], [0, 0]]
    assert v[5]["pos"].tolist() == [[-1, 0], [-1, 1], [0, 1]]
    assert np.allclose(v[5]["w"], [[1, 1, 1], [0.5 / w, 0, 1 / w], [0, 0, 1]], rtol=1e-6)
    # interpolate the rational-quadratic weights at points of the quarter circle: u^2 - v w == 0
    (a, wa), (b, wb), (c, wc) = [(v[5]["pos"][i].astype(np.float64), v[5]["w"][i].astype(np.float64)) for i in range(3)]
    for ang in np.linspace(math.pi / 2 + 0.05, math.pi - 0.05, 9):
        q = np.array([math.cos(ang), math.sin(ang)])
        m = np.array([[a[0], b[0], c[0]], [a[1], b[1], c[1]], [1, 1, 1]])
        lam = np.linalg.solve(m, [q[0], q[1], 1.0])
        u, vv, ww = lam[0] * wa + lam[1] * wb + lam[2] * wc
        assert abs(u * u - vv * ww) < 1e-6
    assert v[2]["pos"].tolist() == [[1, 0], [1, 0], [0, 1], [-1, 0]]   # fan [s, p1, p2, s'] -> strip [0, n-1, 1, n-2] (src/vertex.rs:28-35)
    assert shape.indices()[2].tolist() == [0, 1, 2, 3, 0xFFFF]


def test_cubic_implicit_vanishes_on_the_curve(oracle):
    """§4 invariant 5: Loop-Blinn weights interpolated over the emitted triangles give k^3 - l m n = 0 on the curve and the
    same sign on each side (checked on curve points that fall inside an emitted triangle)."""
    scene = scenes.mixed_fills(200, types=(2,), seed=4)
    P = scene.paths
    checked = 0
    for s in range(60):
        cur = P.start[s].astype(np.float64)
        for seg in P.segments[2][P.type_begin[2, s]:P.type_begin[2, s + 1]].astype(np.float64):
            single = Path(cur)                                      # one cubic per shape: every emitted triangle belongs to it
            single.push_integral_cubic_curve(seg.reshape(3, 2))
            tris = oracle.shape_from_paths([], PathSoA.from_paths([single])).vertices()[4]
            cp = np.array([cur, seg[0:2], seg[2:4], seg[4:6]])
            for t in np.linspace(0.1, 0.9, 9):
                b = np.array([(1 - t) ** 3, 3 * t * (1 - t) ** 2, 3 * t * t * (1 - t), t ** 3])
                q = b @ cp
                for k in range(0, len(tris), 3):
                    a3 = tris["pos"][k:k + 3].astype(np.float64)
                    m = np.vstack([a3.T, np.ones(3)])
                    if abs(np.linalg.det(m)) < 1e-9:
                        continue
                    lam = np.linalg.solve(m, [q[0], q[1], 1.0])
                    if lam.min() < 0.02:
                        continue
                    kk, ll, mm = lam @ tris["w"][k:k + 3].astype(np.float64)
                    scale = max(abs(kk) ** 3, abs(ll * mm), 1e-12)
                    assert abs(kk ** 3 - ll * mm) <= 2e-3 * scale + 1e-6 * (1 + abs(ll) + abs(mm)), (s, t, kk, ll, mm)   # f32 weights
                    checked += 1
            cur = seg[4:6]
    assert checked > 100


def winding_numbers(poly: np.ndarray, width: int, height: int) -> np.ndarray:
    """Winding number of a closed polyline (pixel coordinates) at every pixel centre, by summed signed angles (float64)."""
    ys, xs = np.mgrid[0:height, 0:width]
    px, py = xs + 0.5, ys + 0.5
    total = np.zeros((height, width))
    a = poly
    b = np.roll(poly, -1, axis=0)
    for (ax, ay), (bx, by) in zip(a, b):
        ux, uy, vx, vy = ax - px, ay - py, bx - px, by - py
        total += np.arctan2(ux * vy - uy * vx, ux * vx + uy * vy)
    return np.rint(total / (2 * np.pi)).astype(np.int64)


def flatten(P: PathSoA, s: int, samples: int = 64) -> np.ndarray:
    pts = [P.start[s].astype(np.float64)]
    cursors = [int(P.type_begin[t, s]) for t in range(5)]
    for t in P.segment_types[P.segment_begin[s]:P.segment_begin[s + 1]]:
        seg = P.segments[t][cursors[t]].astype(np.float64)
        cursors[t] += 1
        p0 = pts[-1]
        ts = np.linspace(0, 1, samples + 1)[1:, None]
        if t == 0:
            pts.extend((1 - ts) * p0 + ts * seg[0:2])
        elif t == 1:
            pts.extend((1 - ts) ** 2 * p0 + 2 * ts * (1 - ts) * seg[0:2] + ts ** 2 * seg[2:4])
        elif t == 2:
            pts.extend((1 - ts) ** 3 * p0 + 3 * ts * (1 - ts) ** 2 * seg[0:2] + 3 * ts ** 2 * (1 - ts) * seg[2:4] + ts ** 3 * seg[4:6])
        elif t == 3:
            w = seg[0]
            num = (1 - ts) ** 2 * p0 + 2 * ts * (1 - ts) * w * seg[1:3] + ts ** 2 * seg[3:5]
            pts.extend(num / ((1 - ts) ** 2 + 2 * ts * (1 - ts) * w + ts ** 2))
        else:
            w = seg[0:4]
            bern = np.hstack([(1 - ts) ** 3, 3 * ts * (1 - ts) ** 2, 3 * ts ** 2 * (1 - ts), ts ** 3]) * w
            cps = np.array([p0, seg[4:6], seg[6:8], seg[8:10]])
            pts.extend((bern @ cps) / bern.sum(1, keepdims=True))
    return np.array(pts)


@pytest.mark.parametrize("mirror", [False, True], ids=["counter_clockwise", "clockwise"])
@pytest.mark.parametrize("rational", [False, True], ids=["integral", "all_kinds"])
def test_stencil_equals_winding_number(oracle, rational, mirror):
    """§4 invariant 1, end to end through tessellation AND the raster rules: after Stencil, the winding bits of every sample
    equal the winding number of the path (independent float64 computation on a finely flattened outline), modulo
    2^winding_bits, away from the outline. Pins fan->strip, front/back orientation, the quadratic weights and the top-left rule.

    Cubics (reference quirk C.9): emit_cubic_curve_triangle! turns every Loop-Blinn triangle to one fixed
    orientation (src/fill.rs:127-129) while normalize_implicit_curve_side keeps a fixed side of the direction of travel
    (src/fill.rs:98-114), so the cubic triangles always count in the same direction whatever the sense of the path. With the
    sign conventions that make the join texcoords continuous and the demo's back-face culling show the front, this is
    exact up to a multiple of 2 everywhere (even-odd fills are always right), gives the right non-zero COVERAGE for paths
    traversed clockwise in model coordinates, and overfills the sliver between curve and control polygon for the other sense."""
    scene = scenes.mixed_fills(24, extent=(160, 120), size=(15.0, 45.0), rational=rational, seed=12 + rational, pixels_per_unit=20.0, mirror=mirror)
    cfg = Configuration(winding_counter_bits=4, clip_nesting_counter_bits=4).to_c()
    transforms = scene.transforms()
    compared = exact_shapes = 0
    for s in range(scene.n_shapes):
        shape = oracle.shape_from_paths([], scene.paths, s, s + 1)
        _, stencil, _, _ = oracle.render(cfg, scene.width, scene.height, [shape], [(0, s, s + 1, 0, 0, 0, 0)], transforms, None)
        poly = (flatten(scene.paths, s) + scene.origins[s]) * scene.pixels_per_unit
        want = winding_numbers(poly, scene.width, scene.height)
        ys, xs = np.mgrid[0:scene.height, 0:scene.width]
        c = np.stack([xs + 0.5, ys + 0.5], -1)[:, :, None, :]
        far = ~(np.abs(c - poly[None, None]).sum(-1).min(-1) < 1.5)   # skip samples within ~1 px of the outline
        got = stencil[..., 0].astype(np.int64)
        kinds = set(scene.paths.segment_types[scene.paths.segment_begin[s]:scene.paths.segment_begin[s + 1]].tolist())
        assert np.any(want != 0) and set(np.unique(want)) <= ({0, -1} if mirror else {0, 1})
        assert np.all(((got - want) % 2 == 0)[far]), f"shape {s}: even-odd parity"
        if not kinds & {_abi.CR_SEG_INTEGRAL_CUBIC, _abi.CR_SEG_RATIONAL_CUBIC}:
            assert np.all(((got - want) % 16 == 0)[far]), f"shape {s}: winding number"
            exact_shapes += 1
        elif mirror:
            assert np.all(((got != 0) == (want != 0))[far]), f"shape {s}: non-zero coverage"
        compared += int(far.sum())
    assert compared > 100000 and exact_shapes >= 1


def _distance_to_polyline(poly: np.ndarray, width: int, height: int, ppu: float) -> np.ndarray:
    """Distance (model units) from every pixel centre to an open polyline given in model units; the scene is mapped y-down."""
    ys, xs = np.mgrid[0:height, 0:width]
    c = np.stack([(xs + 0.5) / ppu, (ys + 0.5) / ppu], -1).astype(np.float64)
    best = np.full((height, width), np.inf)
    for a, b in zip(poly[:-1], poly[1:]):
        ab = b - a
        t = np.clip(((c - a) @ ab) / (ab @ ab), 0.0, 1.0)
        best = np.minimum(best, np.linalg.norm(c - (a + t[..., None] * ab), axis=-1))
    return best


def _stroke_coverage(oracle, path, dynamic, width_px, height_px, ppu):
    from contrast_renderer_b200.renderer import orthographic_transform
    soa = PathSoA.from_paths([path])
    shape = oracle.shape_from_paths([dynamic], soa)
    m = orthographic_transform(width_px / ppu, height_px / ppu).reshape(1, 16)
    _, stencil, _, _ = oracle.render(Configuration().to_c(), width_px, height_px, [shape], [(0, 0, 1, 0, 0, 0, 0)], m, None)
    return stencil[..., 0] != 0


@pytest.mark.parametrize("closed", [False, True], ids=["open_round_caps", "closed"])
def test_round_stroke_is_the_offset_region(oracle, closed):
    """Strokes end to end (rows a10-a15, R1, R8): with round joins and round caps the stroked region of a polyline is exactly
    the set of points within width / 2 of it (Minkowski sum with a disc) - an independent ground truth for offset vertices,
    join wedges, cap geometry, the `joint` / `cap` fragment predicates (src/shaders.wgsl:165-300) and the stroke stencil
    state. Pixels whose centre is closer than 0.02 px to the region's boundary are not compared."""
    ppu, w, h, width = 20.0, 240, 180, 0.9
    pts = np.array([[1.5, 1.5], [6.0, 2.0], [3.0, 5.0], [9.5, 4.0], [8.0, 7.5], [2.0, 7.0]])
    path = Path(pts[0], StrokeOptions(width=width, offset=0.0, miter_clip=1.0, closed=closed))
    for q in pts[1:]:
        path.push_line(q)
    covered = _stroke_coverage(oracle, path, DynamicStrokeOptions.Solid(Join.Round, Cap.Round, Cap.Round), w, h, ppu)
    poly = np.vstack([pts, pts[:1]]) if closed else pts
    dist = _distance_to_polyline(poly, w, h, ppu)
    want = dist <= width / 2
    decided = np.abs(dist - width / 2) > 0.02 / ppu
    assert want.sum() > 5000
    wrong = (covered != want) & decided
    assert not wrong.any(), f"{int(wrong.sum())} pixels differ from the offset region, e.g. {np.argwhere(wrong)[:4].tolist()}"


def test_stroked_circle_is_an_annulus(oracle):
    """Curved strokes (curve.rs uniform tangent angle + emit_curve_stroke!, src/stroke.rs:134-168): the closed stroke of a
    circle built from four rational quadratics (src/path.rs:811-813) is the annulus |r - R| <= width / 2. The curve is
    sampled every 0.1 rad, so the strip deviates from the circle by at most R (1 - cos 0.05) = 0.004 units = 0.08 px here:
    pixels closer than 0.15 px to either rim are not compared."""
    ppu, w, h, width, radius = 20.0, 200, 200, 0.8, 3.0
    circle = Path.from_circle([5.0, 5.0], radius, StrokeOptions(width=width, offset=0.0, miter_clip=1.0, closed=True,
                                                               curve_approximation=CurveApproximation.UniformTangentAngle(0.1)))
    covered = _stroke_coverage(oracle, circle, DynamicStrokeOptions.Solid(Join.Round, Cap.Butt, Cap.Butt), w, h, ppu)
    ys, xs = np.mgrid[0:h, 0:w]
    r = np.hypot((xs + 0.5) / ppu - 5.0, (ys + 0.5) / ppu - 5.0)
    want = np.abs(r - radius) <= width / 2
    decided = np.abs(np.abs(r - radius) - width / 2) > 0.15 / ppu
    wrong = (covered != want) & decided
    assert want.sum() > 5500 and not wrong.any(), f"{int(wrong.sum())} pixels differ from the annulus, e.g. {np.argwhere(wrong)[:4].tolist()}"


def test_miter_stroke_of_a_square_is_a_frame(oracle):
    """Miter joins (src/stroke.rs:53-121): the closed stroke of an axis-aligned square is the outer square minus
    the inner square, corners included, when miter_clip admits the full tip (distance w / sqrt 2 <= miter_clip * w)."""
    ppu, w, h, width = 20.0, 200, 200, 1.0
    path = Path([2.5, 2.5], StrokeOptions(width=width, offset=0.0, miter_clip=1.0, closed=True))
    for q in ([7.5, 2.5], [7.5, 7.5], [2.5, 7.5]):
        path.push_line(q)
    covered = _stroke_coverage(oracle, path, DynamicStrokeOptions.Solid(Join.Miter, Cap.Butt, Cap.Butt), w, h, ppu)
    ys, xs = np.mgrid[0:h, 0:w]
    x, y = (xs + 0.5) / ppu, (ys + 0.5) / ppu
    cheb = np.maximum(np.abs(x - 5.0), np.abs(y - 5.0))
    want = (cheb <= 2.5 + width / 2) & (cheb >= 2.5 - width / 2)
    decided = (np.abs(cheb - (2.5 + width / 2)) > 0.02 / ppu) & (np.abs(cheb - (2.5 - width / 2)) > 0.02 / ppu)
    wrong = (covered != want) & decided
    assert want.sum() > 7000 and not wrong.any(), f"{int(wrong.sum())} pixels differ from the frame, e.g. {np.argwhere(wrong)[:4].tolist()}"


def test_dashed_butt_stroke_of_a_line(oracle):
    """Dashes (src/shaders.wgsl:205-231, descriptor packing src/renderer.rs:29-60): along a straight stroke the arc length in
    units of the stroke width, minus the phase, modulo the pattern length selects dash or gap; with Butt caps a dash is an
    exact rectangle. Pattern [gap 1..1.5], [gap 3..4] (pattern length 4 widths), phase 0.25. A dashed stroke has no end-cap
    test (src/shaders.wgsl:276-278 returns before it), so the pattern simply continues over the half-width cap extensions of
    the strip at both ends (src/stroke.rs:270-293,443-462)."""
    ppu, w, h, width = 20.0, 320, 80, 0.5
    x0, x1, yc = 1.0, 15.0, 2.0
    path = Path([x0, yc], StrokeOptions(width=width, offset=0.0, miter_clip=1.0, closed=False))
    path.push_line([x1, yc])
    dyn = DynamicStrokeOptions.Dashed(Join.Bevel, [DashInterval(1.0, 1.5, Cap.Butt, Cap.Butt), DashInterval(3.0, 4.0, Cap.Butt, Cap.Butt)], 0.25)
    covered = _stroke_coverage(oracle, path, dyn, w, h, ppu)
    ys, xs = np.mgrid[0:h, 0:w]
    x, y = (xs + 0.5) / ppu, (ys + 0.5) / ppu
    pos = np.mod((x - x0) / width - 0.25, 4.0)
    in_gap = ((pos > 1.0) & (pos < 1.5)) | ((pos > 3.0) & (pos < 4.0))
    lo, hi = x0 - width / 2, x1 + width / 2
    want = (np.abs(y - yc) <= width / 2) & (x >= lo) & (x <= hi) & ~in_gap
    edges = np.array([0.0, 1.0, 1.5, 3.0, 4.0])
    near_edge = np.min(np.abs(pos[..., None] - edges), -1) < 0.03 / (ppu * width)
    decided = (np.abs(np.abs(y - yc) - width / 2) > 0.02 / ppu) & (np.abs(x - lo) > 0.02 / ppu) & (np.abs(x - hi) > 0.02 / ppu) & ~near_edge
    wrong = (covered != want) & decided
    assert want.sum() > 1500 and (~want & (np.abs(y - yc) <= width / 2) & (x > x0) & (x < x1)).sum() > 500
    assert not wrong.any(), f"{int(wrong.sum())} pixels differ from the dash pattern, e.g. {np.argwhere(wrong)[:4].tolist()}"


def test_color_cover_leaves_no_winding_residue(oracle):
    """§4 invariant 6: after Color the winding bits are zero inside the hull (src/renderer.rs:747-752) and the colour is the
    premultiplied instance colour where the winding was non-zero."""
    scene = scenes.mixed_fills(30, extent=(200, 150), size=(10.0, 50.0), seed=9, pixels_per_unit=20.0)
    scene.colors[:, 3] = 1.0
    cfg = Configuration().to_c()
    shapes = [oracle.shape_from_paths([], scene.paths, i, i + 1) for i in range(scene.n_shapes)]
    cmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in scenes.stencil_cover_commands(scene.n_shapes)]
    color, stencil, _, covered = oracle.render(cfg, scene.width, scene.height, shapes, cmds, scene.transforms(), scene.colors)
    assert (stencil == 0).all() and covered > 1000
    assert set(np.unique(color[..., 3])) <= {0.0, 1.0}


# ------------------------------------------------------------------------------------------- convex_hull.rs
def test_clip_is_an_intersection(oracle):
    """Nested clipping (src/renderer.rs:253-266,692-754): Stencil + Clip to a circle, then Stencil + Color of a rectangle at
    clip depth 1, then Stencil + UnClip: exactly the pixels inside BOTH shapes are coloured (geometric ground truth), a second
    nested clip intersects once more, and all stencil bits are back to zero at the end."""
    ppu, w, h = 20.0, 240, 200
    circle = Path.from_circle([5.0, 5.0], 3.0)
    rect = Path.from_rect([7.0, 5.5], [3.0, 1.5])
    band = Path.from_rect([6.0, 5.0], [0.8, 4.0])
    soa = PathSoA.from_paths([circle, rect, band])
    shapes = [oracle.shape_from_paths([], soa, i, i + 1) for i in range(3)]
    m = np.tile(orthographic_transform(w / ppu, h / ppu), (3, 1))
    colors = np.array([[1, 0, 0, 1], [0, 1, 0, 1], [0, 0, 1, 1]], np.float32)
    S, CLIP, UNCLIP, COLOR = 0, 1, 2, 3
    ys, xs = np.mgrid[0:h, 0:w]
    x, y = (xs + 0.5) / ppu, (ys + 0.5) / ppu
    in_circle = (x - 5.0) ** 2 + (y - 5.0) ** 2 < 9.0
    in_rect = (np.abs(x - 7.0) < 3.0) & (np.abs(y - 5.5) < 1.5)
    in_band = (np.abs(x - 6.0) < 0.8) & (np.abs(y - 5.0) < 4.0)
    near = (np.abs(np.hypot(x - 5.0, y - 5.0) - 3.0) < 0.03 / ppu)   # the conic outline is exact to rounding; lines are exact
    # one clip level
    cmds = [(0, 0, 1, S, 0, 0, 0), (0, 0, 1, CLIP, 1, 0, 0), (1, 1, 2, S, 1, 0, 0), (1, 1, 2, COLOR, 1, 0, 0), (0, 0, 1, S, 0, 0, 0), (0, 0, 1, UNCLIP, 0, 0, 0)]
    color, stencil, _, covered = oracle.render(Configuration().to_c(), w, h, shapes, cmds, m, colors)
    green = color.reshape(h, w, 4)[:, :, 1] > 0.5
    assert not ((green != (in_circle & in_rect)) & ~near).any() and (stencil == 0).all() and covered == int(green.sum()) > 4000
    # two nested clip levels: circle, then band; the rectangle is drawn at depth 2
    cmds = [(0, 0, 1, S, 0, 0, 0), (0, 0, 1, CLIP, 1, 0, 0), (2, 2, 3, S, 1, 0, 0), (2, 2, 3, CLIP, 2, 0, 0),
            (1, 1, 2, S, 2, 0, 0), (1, 1, 2, COLOR, 2, 0, 0),
            (2, 2, 3, S, 1, 0, 0), (2, 2, 3, UNCLIP, 1, 0, 0), (0, 0, 1, S, 0, 0, 0), (0, 0, 1, UNCLIP, 0, 0, 0)]
    color, stencil, _, _ = oracle.render(Configuration().to_c(), w, h, shapes, cmds, m, colors)
    green = color.reshape(h, w, 4)[:, :, 1] > 0.5
    assert not ((green != (in_circle & in_rect & in_band)) & ~near).any() and (stencil == 0).all() and int(green.sum()) > 800


def test_msaa_samples_are_the_webgpu_pattern(oracle):
    """4x MSAA: sample k of a pixel sits at the WebGPU standard offsets (6,2), (14,6), (2,10), (10,14) / 16 and is
    covered iff that point is inside the shape - checked per sample against an analytic polygon test."""
    ppu, w, h = 16.0, 96, 80
    tri = np.array([[0.7, 0.6], [5.3, 1.9], [2.1, 4.4]])
    path = Path.from_polygon(tri)
    shape = oracle.shape_from_paths([], PathSoA.from_paths([path]))
    m = orthographic_transform(w / ppu, h / ppu).reshape(1, 16)
    _, stencil, _, _ = oracle.render(Configuration(msaa_sample_count=4).to_c(), w, h, [shape], [(0, 0, 1, 0, 0, 0, 0)], m, None)
    assert stencil.shape == (h, w, 4)
    offsets = np.array([[6, 2], [14, 6], [2, 10], [10, 14]]) / 16.0
    ys, xs = np.mgrid[0:h, 0:w]
    a, b, c = tri * ppu
    for k, (ox, oy) in enumerate(offsets):
        px, py = xs + ox, ys + oy
        e = [(q[0] - p[0]) * (py - p[1]) - (q[1] - p[1]) * (px - p[0]) for p, q in ((a, b), (b, c), (c, a))]
        inside = ((e[0] > 0) & (e[1] > 0) & (e[2] > 0)) | ((e[0] < 0) & (e[1] < 0) & (e[2] < 0))
        margin = np.min(np.abs(np.array(e)), 0) > 1e-3
        assert not (((stencil[:, :, k] != 0) != inside) & margin).any(), f"sample {k}"
        assert inside.sum() > 1000
    per_pixel = (stencil != 0).sum(-1)
    assert set(np.unique(per_pixel)) == {0, 1, 2, 3, 4}, "edge pixels are partially covered"


def test_perspective_correct_attribute_interpolation(oracle):
    """Instance matrices with a projective row (src/shaders.wgsl:66-74, attributes `@interpolate(perspective, sample)`): a shape
    bounded by a quadratic and a rational quadratic curve is rendered through a homography and compared with the exact image
    of the model-space region (every pixel centre is mapped BACK to model space and tested there). The implicit tests
    u^2 - v <= 0 / u^2 - vw <= 0 only land on the projected curve if (u, v, w) are interpolated perspective-correctly."""
    w, h = 220, 180
    path = Path([1.0, 1.0])
    path.push_integral_quadratic_curve([[4.0, -1.5], [7.0, 1.5]])
    path.push_rational_quadratic_curve(RationalQuadraticCurveSegment(1.8, [[8.5, 5.5], [2.0, 6.0]]))
    soa = PathSoA.from_paths([path])
    shape = oracle.shape_from_paths([], soa)
    # clip = M (x, y, 0, 1): affine part maps [0, 10] x [0, 8] into NDC, plus a projective row w = 1 + 0.06 x + 0.04 y
    m = np.zeros(16, np.float32)
    m[0], m[5], m[10], m[12], m[13], m[15] = 0.21, -0.27, 1.0, -0.95, 0.97, 1.0
    m[3], m[7] = 0.06, 0.04
    m[4], m[1] = 0.02, -0.015
    _, stencil, _, _ = oracle.render(Configuration().to_c(), w, h, [shape], [(0, 0, 1, 0, 0, 0, 0)], m.reshape(1, 16), None)
    ys, xs = np.mgrid[0:h, 0:w]
    nx, ny = 2.0 * (xs + 0.5) / w - 1.0, 1.0 - 2.0 * (ys + 0.5) / h
    md = m.astype(np.float64)
    a11, a12, b1 = md[0] - nx * md[3], md[4] - nx * md[7], nx * md[15] - md[12]
    a21, a22, b2 = md[1] - ny * md[3], md[5] - ny * md[7], ny * md[15] - md[13]
    det = a11 * a22 - a12 * a21
    mx, my = (b1 * a22 - a12 * b2) / det, (a11 * b2 - b1 * a21) / det        # model-space position of every pixel centre
    poly = flatten(soa, 0, samples=400)
    total = np.zeros((h, w))
    for (ax, ay), (bx, by) in zip(poly, np.roll(poly, -1, axis=0)):
        ux, uy, vx, vy = ax - mx, ay - my, bx - mx, by - my
        total += np.arctan2(ux * vy - uy * vx, ux * vx + uy * vy)
    inside = np.rint(total / (2 * np.pi)).astype(np.int64) != 0
    dist = np.min(np.hypot(poly[:, 0][None, None] - mx[..., None], poly[:, 1][None, None] - my[..., None]), -1)
    far = dist > 0.06                                                         # ~1 px in model units at this scale
    assert inside.sum() > 4000 and far.mean() > 0.8
    # a pixel centre that falls exactly on an interior fan edge (1/256 px snapping) between a front- and a back-facing triangle
    # of the folded fan gets the top-left rule's +-1 from one of them only: a measure-zero artefact of stencil-then-cover on
    # any rasteriser, one pixel in this frame
    assert int((((stencil[..., 0] != 0) != inside) & far).sum()) <= 2


def test_frustum_clipping_keeps_exactly_the_part_in_front_of_the_eye(oracle):
    """A projective instance matrix whose w changes sign INSIDE the shape (the plane of the shape crosses the eye plane, as with
    the demo's perspective camera, examples/showcase/main.rs:163-201): WebGPU clips the primitives, so exactly the points of the
    shape with w > 0 are drawn. Every pixel centre is mapped back to model space; it must be covered iff that point lies in
    the shape AND has w > 0 (a back-projected point with w < 0 is the 'ghost' image behind the eye and must stay empty)."""
    w, h = 240, 200
    path = Path([1.0, 1.0])
    path.push_integral_quadratic_curve([[4.0, -1.5], [7.0, 1.5]])
    path.push_line([9.0, 3.0])
    path.push_rational_quadratic_curve(RationalQuadraticCurveSegment(1.8, [[8.5, 5.5], [2.0, 6.0]]))
    soa = PathSoA.from_paths([path])
    shape = oracle.shape_from_paths([], soa)
    m = np.zeros(16, np.float32)
    # w = 1 - 0.16 x + 0.01 y is zero near x = 6.4 (the shape spans x in [1, 9]); clip x = -0.0965 (x - 5.35) puts the visible
    # part on the left of the frame (running off to -infinity as w -> 0+) and the ghost of the part behind the eye on the right
    m[0], m[5], m[10], m[12], m[13], m[15] = -0.0965, -0.05, 1.0, 0.516, 0.17, 1.0
    m[3], m[7] = -0.16, 0.01
    m[4], m[1] = 0.004, -0.003
    _, stencil, _, _ = oracle.render(Configuration().to_c(), w, h, [shape], [(0, 0, 1, 0, 0, 0, 0)], m.reshape(1, 16), None)
    ys, xs = np.mgrid[0:h, 0:w]
    nx, ny = 2.0 * (xs + 0.5) / w - 1.0, 1.0 - 2.0 * (ys + 0.5) / h
    md = m.astype(np.float64)
    a11, a12, b1 = md[0] - nx * md[3], md[4] - nx * md[7], nx * md[15] - md[12]
    a21, a22, b2 = md[1] - ny * md[3], md[5] - ny * md[7], ny * md[15] - md[13]
    det = a11 * a22 - a12 * a21
    mx, my = (b1 * a22 - a12 * b2) / det, (a11 * b2 - b1 * a21) / det
    wc = md[3] * mx + md[7] * my + md[15]
    poly = flatten(soa, 0, samples=400)
    total = np.zeros((h, w))
    for (ax, ay), (bx, by) in zip(poly, np.roll(poly, -1, axis=0)):
        ux, uy, vx, vy = ax - mx, ay - my, bx - mx, by - my
        total += np.arctan2(ux * vy - uy * vx, ux * vx + uy * vy)
    inside = (np.rint(total / (2 * np.pi)).astype(np.int64) != 0) & (wc > 0)
    dist = np.min(np.hypot(poly[:, 0][None, None] - mx[..., None], poly[:, 1][None, None] - my[..., None]), -1)
    px_model = np.abs(wc) / 0.05 * 2.0 / h                                     # ~ one pixel in model units where the pixel looks at
    far = (dist > 3.0 * px_model) & (np.abs(wc) > 1e-3)
    covered = stencil[..., 0] != 0
    ghost = (np.rint(total / (2 * np.pi)).astype(np.int64) != 0) & (wc < 0) & far
    assert inside.sum() > 3000 and ghost.sum() > 200, "the scene must show both a visible part and a ghost region"
    assert int((covered & ghost).sum()) == 0, "nothing behind the eye may be drawn"
    assert int(((covered != inside) & far).sum()) <= 2


def test_andrew_hull_invariants(oracle):
    """§4 invariant 7: convex, clockwise (y up), no three collinear points within 1e-4, and it contains every input point."""
    rng = np.random.default_rng(2)
    for n in (3, 4, 10, 200):
        pts = rng.uniform(-3, 3, (n, 2)).astype(np.float32)
        hull = oracle.andrew(pts).astype(np.float64)
        assert 3 <= len(hull) <= n
        a, b, c = hull, np.roll(hull, -1, 0), np.roll(hull, -2, 0)
        cross = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
        assert np.all(cross < -1e-4)                                           # strictly clockwise turns
        for p in pts.astype(np.float64):
            e = (b[:, 0] - a[:, 0]) * (p[1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (p[0] - a[:, 0])
            assert np.all(e <= 1e-3)                                          # inside or on every clockwise edge
    assert oracle.andrew([[0, 0], [1, 1]]).tolist() == [[0, 0], [1, 1]]       # < 3 points are returned as they are
    square = oracle.andrew([[0, 0], [1, 0], [1, 1], [0, 1], [0.5, 0.5], [0.5, 0.0]])
    assert sorted(map(tuple, square.tolist())) == [(0, 0), (0, 1), (1, 0), (1, 1)]   # interior and collinear points dropped


# ------------------------------------------------------------------------------ renderer.rs descriptor packing
def test_dynamic_stroke_descriptor_packing(oracle):
    """convert_dynamic_stroke_options (src/renderer.rs:29-60): 48 bytes, caps nibbles, count_dashed_join bit layout."""
    from contrast_renderer_b200.path import DashInterval
    p = Path([0, 0], StrokeOptions(width=0.1))
    p.push_line([1, 0])
    soa = PathSoA.from_paths([p])
    solid = DynamicStrokeOptions.Solid(Join.Round, Cap.Out, Cap.Left)
    dashed = DynamicStrokeOptions.Dashed(Join.Bevel, [DashInterval(1.0, 2.0, Cap.Round, Cap.In), DashInterval(3.0, 4.5, Cap.Square, Cap.Right)], 0.25)
    shape = oracle.shape_from_paths([solid, dashed], soa)
    raw = shape.stroke_buffer
    assert len(raw) == 96
    d = np.frombuffer(raw.tobytes(), dtype=np.dtype([("gs", "<f4", 4), ("ge", "<f4", 4), ("caps", "<u4"), ("cdj", "<u4"), ("phase", "<f4"), ("pad", "<u4")]))
    assert d["caps"][0] == (int(Cap.Out) | (int(Cap.Left) << 4)) and d["cdj"][0] == int(Join.Round)
    assert d["cdj"][1] == ((2 - 1) << 3) | 4 | int(Join.Bevel) and d["phase"][1] == np.float32(0.25)
    assert d["gs"][1].tolist() == [1.0, 3.0, 0.0, 0.0] and d["ge"][1].tolist() == [2.0, 4.5, 0.0, 0.0]
    # dash_start of interval i goes to the START-cap nibble of interval i-1 (mod n); dash_end to the END-cap nibble of interval i
    want = (int(Cap.Round) << 8) | (int(Cap.In) << 4) | (int(Cap.Square) << 0) | (int(Cap.Right) << 12)
    assert d["caps"][1] == want
    from oracle.oracle import OracleError
    with pytest.raises(OracleError) as e:
        oracle.shape_from_paths([DynamicStrokeOptions.Dashed(Join.Miter, [DashInterval(i, i + 0.5) for i in range(5)], 0.0)], soa)
    assert e.value.status == _abi.CR_ERR_TOO_MANY_DASH_INTERVALS
    p2 = Path([0, 0], StrokeOptions(width=0.1, dynamic_stroke_options_group=3))
    p2.push_line([1, 0])
    with pytest.raises(OracleError) as e:
        oracle.shape_from_paths([solid], PathSoA.from_paths([p2]))
    assert e.value.status == _abi.CR_ERR_DYNAMIC_STROKE_OPTIONS_INDEX_OUT_OF_BOUNDS


def test_transform_anchor():
    """a translation ends up in the last column (tx, ty, ., 1) of the instance matrix; pixel (0,0) maps to
    NDC (-1, 1) and (W, H) to (1, -1) with the orthographic instance matrix used by every scene."""
    m = orthographic_transform(640, 480).reshape(4, 4).T
    assert np.allclose(m @ [0, 0, 0, 1], [-1, 1, 0, 1]) and np.allclose(m @ [640, 480, 0, 1], [1, -1, 0, 1])
    scene = scenes.mixed_fills(3, extent=(640, 480))
    t = scene.transforms()[1].reshape(4, 4).T
    o = scene.origins[1]
    assert np.allclose(t @ [0, 0, 0, 1], scene.transform().reshape(4, 4).T @ [o[0], o[1], 0, 1], atol=1e-6)


def test_oracle_outputs_are_frozen(oracle):
    """tests/golden/oracle_hashes.json (made by tests/golden/make_oracle_hashes.py): the oracle's bytes for four seeded scenes
    have not changed. Not reference vectors - a guard on the checker itself."""
    import importlib.util
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_oracle_hashes", os.path.join(here, "make_oracle_hashes.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    with open(os.path.join(here, "oracle_hashes.json")) as f:
        want = json.load(f)
    got = module.compute(oracle)
    assert sorted(got) == sorted(want)
    compared = 0
    for name in want:
        if got[name]["inputs"] != want[name]["inputs"]:
            continue   # the generated scene differs on this machine (numpy / CPU): nothing can be said about the oracle
        assert got[name] == want[name], name
        compared += 1
    if compared == 0:
        pytest.skip("no scene reproduced bit-identical inputs on this machine; regenerate tests/golden/oracle_hashes.json")
