"""Seeded synthetic scenes for the configurations named in BASELINE.json, built directly as `PathSoA`
with vectorised numpy so that 10^5..10^6 paths are generated in well under a second.

Every generator is a pure function of its arguments (PCG64 seeded with 0xC0FFEE + config index by default), so the
CPU oracle and the CUDA path — and the GPU box and this container — see bit-identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _abi
from .path import Cap, DashInterval, DynamicStrokeOptions, Join, PathSoA

SEED0 = 0xC0FFEE

# MODEL UNITS. The reference's tolerances are absolute (ERROR_MARGIN = 1e-4, src/error.rs:19, compared against signed
# AREAS in src/fill.rs:127,153 and src/convex_hull.rs:16), so it is only usable with path coordinates of order 1-10 —
# its own demo draws 2.7-unit glyphs and a 5.8-unit rectangle (examples/showcase/main.rs:59-94) and leaves the
# mapping to pixels to the instance matrix. With raw pixel coordinates (areas ~1e4, f32 ulp ~1e-3) the enclosing-
# triangle test of fill.rs:151-156 fails by rounding and the assert at fill.rs:174 panics. The generators therefore
# emit model units and a scene-wide `pixels_per_unit` that goes into the instance transform; scenes with cubic fills
# additionally keep every Shape in LOCAL coordinates around its own origin (areas are differences of products of
# absolute coordinates) and place it with a per-instance translation, as the reference's demo does.


@dataclass
class Scene:
    paths: PathSoA
    shape_path_begin: np.ndarray          # [n_shapes + 1] u32
    dynamic_stroke_options: List[DynamicStrokeOptions]
    width: int
    height: int
    colors: Optional[np.ndarray] = None   # [n_shapes, 4] f32 straight-alpha rgba (one instance per shape)
    name: str = ""
    pixels_per_unit: float = 1.0          # path coordinates are in model units (see MODEL UNITS below)
    origins: Optional[np.ndarray] = None  # [n_shapes, 2] model-space position of each shape's local origin (None: all zero)
    base_transform: Optional[np.ndarray] = None   # overrides the default y-down orthographic mapping (16 floats)

    def transform(self) -> np.ndarray:
        """The instance mat4 (16 floats, column vectors) that maps model units to this scene's framebuffer."""
        if self.base_transform is not None:
            return np.asarray(self.base_transform, np.float32)
        from .renderer import orthographic_transform
        return orthographic_transform(self.width / self.pixels_per_unit, self.height / self.pixels_per_unit)

    def transforms(self) -> np.ndarray:
        """[n_shapes, 16]: instance i = scene transform * translate(origins[i]) (column vectors, src/shaders.wgsl:13-27)."""
        m = np.tile(self.transform(), (self.n_shapes, 1))
        if self.origins is not None:
            o = np.asarray(self.origins, np.float64)
            m[:, 12] = (m[:, 0].astype(np.float64) * o[:, 0] + m[:, 12]).astype(np.float32)
            m[:, 13] = (m[:, 5].astype(np.float64) * o[:, 1] + m[:, 13]).astype(np.float32)
        return np.ascontiguousarray(m, np.float32)

    @property
    def n_shapes(self) -> int:
        return len(self.shape_path_begin) - 1


def _assemble(start: np.ndarray, seg_counts: np.ndarray, seg_types: np.ndarray, payload: List[np.ndarray], stroke: np.ndarray) -> PathSoA:
    """payload[t]: rows (in global segment order) of the segments whose type is t."""
    n = len(seg_counts)
    segment_begin = np.zeros(n + 1, np.uint32)
    np.cumsum(seg_counts, out=segment_begin[1:])
    seg_path = np.repeat(np.arange(n), seg_counts)
    type_begin = np.zeros((5, n + 1), np.uint32)
    for t in range(5):
        per_path = np.bincount(seg_path[seg_types == t], minlength=n)
        np.cumsum(per_path, out=type_begin[t, 1:])
    segments = [np.ascontiguousarray(payload[t], dtype=np.float32).reshape(-1, _abi.SEGMENT_FLOATS[t]) for t in range(5)]
    # SafeFloat canonicalisation of the inputs (src/safe_float.rs:46-49): -0.0 -> +0.0
    segments = [s + np.float32(0.0) for s in segments]
    return PathSoA(np.ascontiguousarray(start, np.float32) + np.float32(0.0), segment_begin, np.ascontiguousarray(seg_types, np.uint8), type_begin, segments,
                   stroke)


def _stroke_records(n: int, width, offset=0.0, miter_clip=4.0, closed=True, group=0, angle_step: Optional[float] = 0.1, steps: int = 0) -> np.ndarray:
    rec = np.zeros(n, PathSoA.STROKE_DTYPE)
    rec["width"] = width
    rec["offset"] = offset
    rec["miter_clip"] = miter_clip
    flags = _abi.CR_STROKE_FLAG_STROKED | (_abi.CR_STROKE_FLAG_CLOSED if closed else 0)
    if angle_step is not None:
        flags |= _abi.CR_STROKE_FLAG_UNIFORM_TANGENT_ANGLE
        rec["approx"] = np.float32(angle_step).view(np.uint32)
    else:
        rec["approx"] = steps
    rec["flags"] = flags
    rec["group"] = group
    return rec


def _blob_geometry(rng: np.random.Generator, centre: np.ndarray, radius: np.ndarray, seg_counts: np.ndarray):
    """Star-convex closed outlines: anchors at increasing angles around `centre`. Returns per-segment (a, b, na, nb):
    endpoints and outward unit normals at the endpoints, in global segment order, plus the path starts."""
    n = len(seg_counts)
    total = int(seg_counts.sum())
    seg_path = np.repeat(np.arange(n), seg_counts)
    first = np.zeros(n + 1, np.int64)
    np.cumsum(seg_counts, out=first[1:])
    local = np.arange(total) - first[seg_path]
    k = seg_counts[seg_path].astype(np.float64)
    phase = rng.uniform(0.0, 2.0 * np.pi, n)[seg_path]
    jitter = rng.uniform(-0.25, 0.25, total)
    ang = phase + (local + jitter) * (2.0 * np.pi / k)
    rad = radius[seg_path] * rng.uniform(0.55, 1.0, total)
    pts = centre[seg_path] + np.stack([np.cos(ang), np.sin(ang)], 1) * rad[:, None]
    nxt = np.where(local + 1 == seg_counts[seg_path], first[seg_path], np.arange(total) + 1)
    a, b = pts, pts[nxt]
    out_a = np.stack([np.cos(ang), np.sin(ang)], 1)
    out_b = out_a[nxt]
    return a, b, out_a, out_b, pts[first[:-1]], seg_path


def _shape_layout(rng: np.random.Generator, n_paths: int, paths_per_shape: int, extent, ppu: float, spread: np.ndarray):
    """Shapes of `paths_per_shape` consecutive paths; returns (shape_path_begin, shape origins in model units, path
    centres in the LOCAL coordinates of their shape: zero for single-path shapes, else within +-1.5 * spread)."""
    begin = np.arange(0, n_paths + 1, paths_per_shape, dtype=np.uint32)
    if begin[-1] != n_paths:
        begin = np.append(begin, np.uint32(n_paths))
    n_shapes = len(begin) - 1
    origins = np.stack([rng.uniform(0, extent[0], n_shapes), rng.uniform(0, extent[1], n_shapes)], 1) / ppu
    local = rng.uniform(-1.5, 1.5, (n_paths, 2)) * np.asarray(spread, np.float64).reshape(-1, 1)
    if paths_per_shape == 1:
        local[:] = 0.0
    return begin, origins, local


# ------------------------------------------------------------------------------------------------------ config 1
def _rotate(v: np.ndarray, angle: np.ndarray) -> np.ndarray:
    c, s = np.cos(angle)[:, None], np.sin(angle)[:, None]
    return np.concatenate([v[:, :1] * c - v[:, 1:] * s, v[:, :1] * s + v[:, 1:] * c], 1)


def closed_cubic_strokes(n_paths: int = 1000, seed: int = SEED0 + 1, extent: Tuple[int, int] = (1920, 1080), paths_per_shape: int = 1,
                         pixels_per_unit: float = 40.0, angle_step: float = 0.1) -> Scene:
    """BASELINE config 1: closed paths of 4 integral cubics, width U[1,8] px, miter clip 4, UniformTangentAngle(0.1).
    Handles are rotated off the blob tangent by U[-0.7, 0.7] rad, so most anchors are corners (miter joins)."""
    rng = np.random.default_rng(seed)
    ppu = float(pixels_per_unit)
    seg_counts = np.full(n_paths, 4, np.int64)
    radius = rng.uniform(20.0, 200.0, n_paths) / ppu
    begin, origins, centre = _shape_layout(rng, n_paths, paths_per_shape, extent, ppu, spread=radius)
    a, b, na, nb, start, _ = _blob_geometry(rng, centre, radius, seg_counts)
    chord = b - a
    tang = _rotate(np.stack([-na[:, 1], na[:, 0]], 1), rng.uniform(-0.7, 0.7, len(a)))   # counter-clockwise tangent at a, perturbed
    tang_b = _rotate(np.stack([-nb[:, 1], nb[:, 0]], 1), rng.uniform(-0.7, 0.7, len(a)))
    clen = np.linalg.norm(chord, axis=1, keepdims=True)
    h0 = rng.uniform(0.2, 0.6, (len(a), 1)) * clen
    h1 = rng.uniform(0.2, 0.6, (len(a), 1)) * clen
    cubic = np.concatenate([a + tang * h0, b - tang_b * h1, b], 1)
    types = np.full(len(a), _abi.CR_SEG_INTEGRAL_CUBIC, np.uint8)
    payload = [np.zeros((0, 2)), np.zeros((0, 4)), cubic, np.zeros((0, 5)), np.zeros((0, 10))]
    stroke = _stroke_records(n_paths, (rng.uniform(1.0, 8.0, n_paths) / ppu).astype(np.float32), angle_step=angle_step)
    soa = _assemble(start, seg_counts, types, payload, stroke)
    n_shapes = len(begin) - 1
    colors = np.concatenate([rng.uniform(0, 1, (n_shapes, 3)), np.ones((n_shapes, 1))], 1).astype(np.float32)
    return Scene(soa, begin, [DynamicStrokeOptions.Solid(Join.Miter, Cap.Butt, Cap.Butt)], extent[0], extent[1], colors, "closed_cubic_strokes", ppu,
                 origins)


def _det3(a, b, c):
    return (a[:, 0] * (b[:, 1] * c[:, 2] - b[:, 2] * c[:, 1]) - a[:, 1] * (b[:, 0] * c[:, 2] - b[:, 2] * c[:, 0])
            + a[:, 2] * (b[:, 0] * c[:, 1] - b[:, 1] * c[:, 0]))


def cubic_fill_is_safe(points: np.ndarray, weights: Optional[np.ndarray] = None) -> np.ndarray:
    """Which cubic segments the reference's FillBuilder can digest (float64 screening of the generator's output).

    points: [n, 4, 2] control points including the start; weights: [n, 4] or None (integral).
    The reference panics (assert_eq!/assert_ne! at src/fill.rs:174,178) or mis-triangulates when
      * the control quadrilateral is (nearly) degenerate, so the signs of the four sub-triangle areas are noise;
      * a loop's double point has exactly one parameter in (0, 1) that lies close to 0 or 1: split_curve_at!
        (src/fill.rs:206-216,232-241) then produces a sliver quadrilateral with noise signs;
      * a RATIONAL cubic has a concave control quadrilateral: the areas are computed from weighted points
        (the `.signum()` normalisation is commented out at src/fill.rs:143), so the enclosing-triangle identity
        of src/fill.rs:151-156 no longer holds.
    Generators demote such segments to lower-order ones; parity tests additionally check every scene with the oracle."""
    n = len(points)
    w = np.ones((n, 4)) if weights is None else np.asarray(weights, np.float64)
    h = np.concatenate([w[:, :, None], points * w[:, :, None]], 2)          # (w, wx, wy)
    pb = [h[:, 0], -3 * h[:, 0] + 3 * h[:, 1], 3 * h[:, 0] - 6 * h[:, 1] + 3 * h[:, 2], -h[:, 0] + 3 * h[:, 1] - 3 * h[:, 2] + h[:, 3]]
    d = np.stack([_det3(pb[1], pb[2], pb[3]), -_det3(pb[0], pb[2], pb[3]), _det3(pb[0], pb[1], pb[3]), -_det3(pb[0], pb[1], pb[2])], 1)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-300)
    integral = np.abs(d[:, 0]) <= 1e-4
    with np.errstate(all="ignore"):
        disc_i = 3 * d[:, 2] ** 2 - 4 * d[:, 1] * d[:, 3]
        sq = np.sqrt(np.maximum(-disc_i, 0))
        ri = np.stack([(d[:, 2] + sq) / (2 * d[:, 1]), (d[:, 2] - sq) / (2 * d[:, 1])], 1)
        c0, c1, c2 = d[:, 1] * d[:, 3] - d[:, 2] ** 2, d[:, 1] * d[:, 2] - d[:, 0] * d[:, 3], d[:, 0] * d[:, 2] - d[:, 1] ** 2
        disc_r = c1 * c1 - 4 * c2 * c0
        sr = np.sqrt(np.maximum(disc_r, 0))
        rr = np.stack([(-c1 - sr) / (2 * c2), (-c1 + sr) / (2 * c2)], 1)
    loop = np.where(integral, (disc_i < 0) & (np.abs(d[:, 1]) > 1e-4), disc_r > 0)
    roots = np.where(integral[:, None], ri, rr)
    roots = np.where(np.isfinite(roots), roots, 1e9)
    near = ((np.abs(roots) < 0.2) | (np.abs(roots - 1) < 0.2)).any(1)
    # rational loops also carry ONE real inflection point (root of -d3 + 3 d2 t - 3 d1 t^2 + d0 t^3); when it is the only
    # root in (0, 1) the reference splits there (find_double_point_issue, src/fill.rs:14-32) and the split
    # quadrilaterals have a zero-area sub-triangle by construction => noise signs => panic
    bad_inflection = np.zeros(n, bool)
    if weights is not None:
        for i in np.nonzero(loop & ~integral)[0]:
            r = np.roots([d[i, 0], -3 * d[i, 1], 3 * d[i, 2], -d[i, 3]])
            r = r[np.abs(r.imag) < 1e-9].real
            bad_inflection[i] = bool(((r > -0.05) & (r < 1.05)).any())
    bad_loop = loop & (near | bad_inflection)
    # nearly singular classification (cusp-like or almost-quadratic) is where f32 and f64 disagree: stay clear of it
    fragile = np.where(integral, np.abs(disc_i) < 1e-3, np.abs(disc_r) < 1e-3) | (~integral & (np.abs(d[:, 0]) < 1e-3))
    u = np.concatenate([np.ones((n, 4, 1)), points], 2)
    ua = np.stack([_det3(u[:, 1], u[:, 2], u[:, 3]), _det3(u[:, 0], u[:, 2], u[:, 3]), _det3(u[:, 0], u[:, 1], u[:, 3]), _det3(u[:, 0], u[:, 1], u[:, 2])], 1)
    wa = ua * np.stack([w[:, 1] * w[:, 2] * w[:, 3], w[:, 0] * w[:, 2] * w[:, 3], w[:, 0] * w[:, 1] * w[:, 3], w[:, 0] * w[:, 1] * w[:, 2]], 1)
    degenerate = np.abs(wa).min(1) < 1e-2
    concave = 2 * np.abs(ua).max(1) > 0.97 * np.abs(ua).sum(1)
    bad = bad_loop | fragile | degenerate
    if weights is not None:
        bad |= concave
    return ~bad


# ------------------------------------------------------------------------------------------------------ config 2
def mixed_fills(n_paths: int = 10000, seed: int = SEED0 + 2, extent: Tuple[int, int] = (1920, 1080), size: Tuple[float, float] = (10.0, 150.0),
                seg_range: Tuple[int, int] = (3, 12), types=(0, 1, 2), paths_per_shape: int = 1, rational: bool = False,
                pixels_per_unit: float = 40.0, mirror: bool = False) -> Scene:
    """BASELINE config 2: filled closed paths of 3..12 segments drawn from {line, integral quadratic, integral cubic}
    (or, with rational=True, also the rational kinds with weights U[0.5, 2]). Outlines run with increasing angle in model
    space; mirror=True negates the local x coordinates, i.e. the same outlines traversed in the opposite sense."""
    rng = np.random.default_rng(seed)
    ppu = float(pixels_per_unit)
    seg_counts = rng.integers(seg_range[0], seg_range[1] + 1, n_paths).astype(np.int64)
    radius = rng.uniform(size[0], size[1], n_paths) / ppu
    begin, origins, centre = _shape_layout(rng, n_paths, paths_per_shape, extent, ppu, spread=radius)
    a, b, na, nb, start, _ = _blob_geometry(rng, centre, radius, seg_counts)
    total = len(a)
    kinds = np.asarray(types if not rational else (0, 1, 2, 3, 4), np.uint8)
    seg_types = kinds[rng.integers(0, len(kinds), total)]
    if mirror:
        flip = np.array([-1.0, 1.0])
        a, b, na, nb, start = a * flip, b * flip, na * flip, nb * flip, start * flip
    chord = b - a
    clen = np.linalg.norm(chord, axis=1, keepdims=True)
    mid = 0.5 * (a + b)
    bulge = rng.uniform(-0.15, 0.45, (total, 1)) * clen
    nmid = na + nb
    nmid /= np.maximum(np.linalg.norm(nmid, axis=1, keepdims=True), 1e-9)
    quad_c = mid + nmid * bulge
    turn = -1.0 if mirror else 1.0   # tangent along the direction of travel
    tang = np.stack([-na[:, 1], na[:, 0]], 1) * turn
    tang_b = np.stack([-nb[:, 1], nb[:, 0]], 1) * turn
    h0 = rng.uniform(0.15, 0.7, (total, 1)) * clen
    h1 = rng.uniform(0.15, 0.7, (total, 1)) * clen
    s0 = rng.uniform(-0.3, 0.5, (total, 1)) * clen
    s1 = rng.uniform(-0.3, 0.5, (total, 1)) * clen
    c1 = a + tang * h0 + na * s0
    c2 = b - tang_b * h1 + nb * s1
    wq = rng.uniform(0.5, 2.0, (total, 1))
    wc = rng.uniform(0.5, 2.0, (total, 4))
    # cubics the reference's fill builder cannot digest become quadratics of the same kind (see cubic_fill_is_safe)
    quad_pts = np.stack([a, c1, c2, b], 1)
    ic, rc = seg_types == _abi.CR_SEG_INTEGRAL_CUBIC, seg_types == _abi.CR_SEG_RATIONAL_CUBIC
    seg_types = np.where(ic & ~cubic_fill_is_safe(quad_pts), _abi.CR_SEG_INTEGRAL_QUADRATIC, seg_types)
    seg_types = np.where(rc & ~cubic_fill_is_safe(quad_pts, wc), _abi.CR_SEG_RATIONAL_QUADRATIC, seg_types).astype(np.uint8)
    sel = [seg_types == t for t in range(5)]
    payload = [b[sel[0]], np.concatenate([quad_c, b], 1)[sel[1]], np.concatenate([c1, c2, b], 1)[sel[2]], np.concatenate([wq, quad_c, b], 1)[sel[3]],
               np.concatenate([wc, c1, c2, b], 1)[sel[4]]]
    stroke = np.zeros(n_paths, PathSoA.STROKE_DTYPE)
    soa = _assemble(start, seg_counts, seg_types, payload, stroke)
    n_shapes = len(begin) - 1
    colors = np.concatenate([rng.uniform(0, 1, (n_shapes, 3)), np.ones((n_shapes, 1))], 1).astype(np.float32)
    return Scene(soa, begin, [], extent[0], extent[1], colors, "mixed_fills", ppu, origins)


# ------------------------------------------------------------------------------------------------------ config 3
def glyph_like_fills(n_glyphs: int = 100000, seed: int = SEED0 + 3, extent: Tuple[int, int] = (3840, 2160), em: float = 20.0,
                     glyphs_per_shape: int = 400, pixels_per_unit: float = 10.0) -> Scene:
    """BASELINE config 3 stand-in until the text front-end lands: glyph-sized closed contours of lines and
    integral quadratics (TrueType outlines are quadratic) with the measured OpenSans statistics — 1.44 contours and 20.9
    outline points per glyph — laid out on text lines; one Shape per run of `glyphs_per_shape` glyphs,
    like `paths_of_text` output chunked per line. Inner contours are reversed (counter shapes), so non-zero and
    even-odd fills differ from a plain union."""
    rng = np.random.default_rng(seed)
    ppu = float(pixels_per_unit)
    has_inner = rng.uniform(0, 1, n_glyphs) < 0.44
    n_paths = int(n_glyphs + has_inner.sum())
    glyph_of_path = np.concatenate([np.arange(n_glyphs), np.nonzero(has_inner)[0]])
    inner = np.concatenate([np.zeros(n_glyphs, bool), np.ones(int(has_inner.sum()), bool)])
    order = np.lexsort((inner, glyph_of_path))   # outer contour first, then its inner contour
    glyph_of_path, inner = glyph_of_path[order], inner[order]
    advance = 0.6 * em
    line_height = 1.36 * em   # OpenSans: (ascender - descender + gap) / unitsPerEm = 2789 / 2048
    per_line = max(1, int((extent[0] - 2 * em) // advance))
    col, row = glyph_of_path % per_line, glyph_of_path // per_line
    n_rows_fit = max(1, int((extent[1] - em) // line_height))
    centre = np.stack([em + (col + 0.5) * advance, em * 0.8 + (row % n_rows_fit) * line_height], 1).astype(np.float64) / ppu
    radius = np.where(inner, 0.16 * em, 0.36 * em) * rng.uniform(0.8, 1.0, n_paths) / ppu
    seg_counts = np.where(inner, rng.integers(4, 9, n_paths), rng.integers(8, 15, n_paths)).astype(np.int64)
    a, b, na, nb, start, seg_path = _blob_geometry(rng, centre, radius, seg_counts)
    total = len(a)
    seg_types = np.where(rng.uniform(0, 1, total) < 0.55, _abi.CR_SEG_INTEGRAL_QUADRATIC, _abi.CR_SEG_LINE).astype(np.uint8)
    clen = np.linalg.norm(b - a, axis=1, keepdims=True)
    nmid = na + nb
    nmid /= np.maximum(np.linalg.norm(nmid, axis=1, keepdims=True), 1e-9)
    quad_c = 0.5 * (a + b) + nmid * rng.uniform(-0.1, 0.4, (total, 1)) * clen
    # reverse inner contours: walk the same outline backwards (src/path.rs:445-488 `reverse` does this in the showcase)
    rev = inner[seg_path]
    if rev.any():
        first = np.zeros(n_paths + 1, np.int64)
        np.cumsum(seg_counts, out=first[1:])
        local = np.arange(total) - first[seg_path]
        mirror = first[seg_path] + (seg_counts[seg_path] - 1 - local)
        src = np.where(rev, mirror, np.arange(total))
        a2, b2 = np.where(rev[:, None], b[src], a[src]), np.where(rev[:, None], a[src], b[src])
        quad_c, seg_types = quad_c[src], seg_types[src]
        a, b = a2, b2
        start = a[first[:-1]]
    sel = [seg_types == t for t in range(5)]
    payload = [b[sel[0]], np.concatenate([quad_c, b], 1)[sel[1]], np.zeros((0, 6)), np.zeros((0, 5)), np.zeros((0, 10))]
    stroke = np.zeros(n_paths, PathSoA.STROKE_DTYPE)
    soa = _assemble(start, seg_counts, seg_types, payload, stroke)
    # shapes = runs of glyphs_per_shape glyphs (paths of one glyph stay together)
    glyph_first_path = np.searchsorted(glyph_of_path, np.arange(0, n_glyphs, glyphs_per_shape))
    begin = np.append(glyph_first_path, n_paths).astype(np.uint32)
    n_shapes = len(begin) - 1
    colors = np.concatenate([rng.uniform(0, 0.8, (n_shapes, 3)), np.ones((n_shapes, 1))], 1).astype(np.float32)
    return Scene(soa, begin, [], extent[0], extent[1], colors, "glyph_like_fills", ppu)


def text_glyphs(n_glyphs: int = 100000, seed: int = SEED0 + 3, extent: Tuple[int, int] = (3840, 2160), size_px: float = 12.0, chars_per_line: int = 640,
                pixels_per_unit: float = 60.0, glyphs_per_shape: int = 160) -> Scene:
    """BASELINE config 3 through the text front-end (src/text.rs): `n_glyphs` printable ASCII characters (U+0021..U+007E,
    64-bit LCG), a newline every `chars_per_line`, laid out by `paths_of_text` semantics with OpenSans outlines
    (tests/golden/opensans_ascii.npz, extracted from the reference's demo font), one Shape per run of `glyphs_per_shape`
    characters (a phrase; the reference's demo puts one string into one Shape). Text is y-up and centred on the origin like
    the reference's demo; the instance matrix maps it to the centre of the target without mirroring. 100 000 glyphs at 12 px
    fill 157 lines of a 3840x2160 target. pixels_per_unit = 60 keeps the model coordinates of the whole page within +-30
    units, where the reference's absolute 1e-4 hull tolerance (src/convex_hull.rs:16) still removes collinear baseline
    points; at +-260 units f32 noise exceeds the tolerance and every line's hull keeps ~1000 vertices."""
    import os
    from .text import Alignment, FixtureFace, Layout, Orientation, text_to_soa
    fixture = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "opensans_ascii.npz")
    face = FixtureFace(dict(np.load(fixture)))
    state = np.uint64(seed)
    codes = np.empty(n_glyphs, np.int64)
    with np.errstate(over="ignore"):
        for i in range(n_glyphs):   # Knuth MMIX LCG, high bits
            state = state * np.uint64(6364136223846793005) + np.uint64(1442695040888963407)
            codes[i] = 33 + int(state >> np.uint64(33)) % 94
    chars = [chr(c) for c in codes]
    lines = ["".join(chars[i:i + chars_per_line]) for i in range(0, n_glyphs, chars_per_line)]
    text = "\n".join(lines)
    ppu = float(pixels_per_unit)
    layout = Layout(size_px / ppu, Orientation.LeftToRight, Alignment.Center, Alignment.Center)
    soa, occ_first_path = text_to_soa(face, layout, text)
    shape_first_glyph = np.arange(0, n_glyphs, glyphs_per_shape)
    begin = np.append(occ_first_path[shape_first_glyph], soa.n_paths).astype(np.uint32)
    n_shapes = len(begin) - 1
    rng = np.random.default_rng(seed)
    colors = np.concatenate([rng.uniform(0, 0.8, (n_shapes, 3)), np.ones((n_shapes, 1))], 1).astype(np.float32)
    m = np.zeros(16, np.float32)
    m[0], m[5], m[10], m[15] = 2.0 * ppu / extent[0], 2.0 * ppu / extent[1], 1.0, 1.0
    return Scene(soa, begin, [], extent[0], extent[1], colors, "text_glyphs", ppu, None, m)


# ------------------------------------------------------------------------------------------------------ config 5
def dashed_rational_strokes(n_paths: int = 1000000, seed: int = SEED0 + 5, extent: Tuple[int, int] = (7680, 4320), paths_per_shape: int = 1000,
                            angle_step: float = 0.2, pixels_per_unit: float = 10.0) -> Scene:
    """BASELINE config 5: open paths of 2 rational cubics (weights U[0.5,2]), width U[1,4], round joins and caps,
    two-interval dash pattern, UniformTangentAngle(0.2). A Shape is the `paths_per_shape` paths of one cell of a grid over
    the target, in LOCAL coordinates around the cell centre, placed by its instance translation (MODEL UNITS, top of this
    file): with absolute coordinates of hundreds of units the reference's hull tolerance drowns in f32 noise and every
    Shape's "hull" keeps ~1500 vertices in overlapping slivers (measured: 13 (tile, primitive) pairs per tile per cover)."""
    rng = np.random.default_rng(seed)
    ppu = float(pixels_per_unit)
    seg_counts = np.full(n_paths, 2, np.int64)
    n_cells = (n_paths + paths_per_shape - 1) // paths_per_shape
    gx = max(1, int(round(np.sqrt(n_cells * extent[0] / extent[1]))))
    gy = (n_cells + gx - 1) // gx
    cell_w, cell_h = extent[0] / gx, extent[1] / gy
    cell = np.arange(n_cells)
    origins = np.stack([(cell % gx + 0.5) * cell_w, (cell // gx + 0.5) * cell_h], 1) / ppu
    p0 = np.stack([rng.uniform(-0.5 * cell_w, 0.5 * cell_w, n_paths), rng.uniform(-0.5 * cell_h, 0.5 * cell_h, n_paths)], 1) / ppu
    step = rng.uniform(8.0, 30.0, (n_paths, 1)) / ppu
    direction = rng.uniform(0, 2 * np.pi, n_paths)
    d = np.stack([np.cos(direction), np.sin(direction)], 1)
    nrm = np.stack([-d[:, 1], d[:, 0]], 1)

    def wiggle(scale):
        return nrm * rng.uniform(-1.0, 1.0, (n_paths, 1)) * step * scale

    q1 = p0 + d * step + wiggle(0.8)
    q2 = p0 + d * step * 2 + wiggle(0.8)
    q3 = p0 + d * step * 3 + wiggle(0.3)
    turn = rng.uniform(-0.9, 0.9, n_paths)
    d2 = np.stack([np.cos(direction + turn), np.sin(direction + turn)], 1)
    n2 = np.stack([-d2[:, 1], d2[:, 0]], 1)
    r1 = q3 + d2 * step + n2 * rng.uniform(-1.0, 1.0, (n_paths, 1)) * step * 0.8
    r2 = q3 + d2 * step * 2 + n2 * rng.uniform(-1.0, 1.0, (n_paths, 1)) * step * 0.8
    r3 = q3 + d2 * step * 3
    w = rng.uniform(0.5, 2.0, (n_paths, 2, 4))
    seg = np.stack([np.concatenate([w[:, 0], q1, q2, q3], 1), np.concatenate([w[:, 1], r1, r2, r3], 1)], 1).reshape(-1, 10)
    types = np.full(2 * n_paths, _abi.CR_SEG_RATIONAL_CUBIC, np.uint8)
    payload = [np.zeros((0, 2)), np.zeros((0, 4)), np.zeros((0, 6)), np.zeros((0, 5)), seg]
    stroke = _stroke_records(n_paths, (rng.uniform(1.0, 4.0, n_paths) / ppu).astype(np.float32), miter_clip=1.0, closed=False, angle_step=angle_step)
    soa = _assemble(p0, seg_counts, types, payload, stroke)
    begin = np.arange(0, n_paths + 1, paths_per_shape, dtype=np.uint32)
    if begin[-1] != n_paths:
        begin = np.append(begin, np.uint32(n_paths))
    n_shapes = len(begin) - 1
    colors = np.concatenate([rng.uniform(0, 1, (n_shapes, 3)), np.full((n_shapes, 1), 0.8)], 1).astype(np.float32)
    dso = DynamicStrokeOptions.Dashed(Join.Round, [DashInterval(2.0, 3.0, Cap.Round, Cap.Round), DashInterval(5.0, 5.5, Cap.Round, Cap.Round)],
                                      float(rng.uniform(0, 6)))
    return Scene(soa, begin, [dso], extent[0], extent[1], colors, "dashed_rational_strokes", ppu, origins[:n_shapes])


def stencil_cover_commands(n_shapes: int) -> np.ndarray:
    """One Stencil + one Color per shape, shape i drawn with instance i (examples/showcase/main.rs:236-249)."""
    cmds = np.zeros((2 * n_shapes, 4), np.uint32)
    idx = np.arange(n_shapes, dtype=np.uint32)
    cmds[0::2] = np.stack([idx, idx, idx + 1, np.zeros_like(idx)], 1)
    cmds[1::2] = np.stack([idx, idx, idx + 1, np.full_like(idx, 3)], 1)
    return cmds


# ------------------------------------------------------------------------------------------------------ config 4
OP_STENCIL, OP_CLIP, OP_UNCLIP, OP_COLOR, OP_SAVE_ALPHA, OP_SCALE_ALPHA, OP_RESTORE_ALPHA = range(7)   # RenderOperation, src/renderer.rs:144-174


@dataclass
class ScriptedScene:
    """A scene whose draws carry pass state: `script` rows are (shape, instance_begin, instance_end, operation,
    clip_depth, save_alpha_layer, restore_alpha_layer), the state being what `set_clip_depth` /
    `save_alpha_context` / `restore_alpha_context` were last called with (src/renderer.rs:253-266,932-985)."""
    paths: PathSoA
    shape_path_begin: np.ndarray
    dynamic_stroke_options: List[DynamicStrokeOptions]
    width: int
    height: int
    transforms: np.ndarray     # [n_instances, 16]
    colors: np.ndarray         # [n_instances, 4]
    script: np.ndarray         # [n_commands, 7] u32
    name: str = ""
    alpha_layer_count: int = 2

    @property
    def n_shapes(self) -> int:
        return len(self.shape_path_begin) - 1

    def oracle_commands(self):
        return [tuple(int(v) for v in row) for row in self.script]

    def record(self, render_pass, batch, one_call: bool = True) -> None:
        """Replays the script into a RenderPass: one bulk call (`cr_pass_render_script`), or — one_call=False — the individual
        state calls where the state changes with bulk recording in between (the two are equivalent by definition)."""
        if one_call:
            render_pass.render_script(batch, self.script)
            return
        script = self.script
        state = (None, None, None)
        run_start = 0
        for i in range(len(script) + 1):
            new_state = tuple(int(v) for v in script[i, 4:7]) if i < len(script) else None
            if new_state != state:
                if i > run_start:
                    render_pass.render_batch(batch, script[run_start:i, 0:4])
                run_start = i
                if new_state is not None:
                    if new_state[0] != state[0]:
                        render_pass.set_clip_depth(new_state[0])
                    if new_state[1] != state[1]:
                        render_pass.save_alpha_context(new_state[1])
                    if new_state[2] != state[2]:
                        render_pass.restore_alpha_context(new_state[2])
                    state = new_state


def tiger_like(n_instances: int = 1000, seed: int = SEED0 + 4, extent: Tuple[int, int] = (3840, 2160), paths_per_shape: int = 10,
               instance_px: Tuple[float, float] = (60.0, 220.0)) -> ScriptedScene:
    """BASELINE config 4, "tiger-style": `n_instances` placed copies (translation, rotation, scale) of one group of
    24 Shapes x `paths_per_shape` paths built with the path constructors (src/path.rs:639-815: ellipses, circles, rounded
    rectangles, elliptical-arc wedges -> rational quadratics; the same degree-elevated -> rational cubics). Every copy
    runs three nested clips and two nested opacity groups (src/renderer.rs:253-266). Each (copy, Shape) pair is its own
    instance (transform + colour), 24 per copy."""
    from .path import Path
    rng = np.random.default_rng(seed)
    n_roles = 24

    def decorations(cx, cy, rx, ry, count, cubic):
        """`count` small conic outlines inside the box (cx +- rx, cy +- ry)."""
        out = []
        for k in range(count):
            x, y = cx + rng.uniform(-0.6, 0.6) * rx, cy + rng.uniform(-0.6, 0.6) * ry
            r = rng.uniform(0.12, 0.3) * min(rx, ry)
            kind = int(rng.integers(0, 4))
            if kind == 0:
                p = Path.from_circle([x, y], r)
            elif kind == 1:
                p = Path.from_ellipse([x, y], [r * 1.4, r * 0.7])
            elif kind == 2:
                p = Path.from_rounded_rect([x, y], [r * 1.3, r * 0.9], r * 0.35)
            else:   # pie wedge: centre -> arc start -> elliptical arc -> implicit closing line
                a0, a1 = rng.uniform(0, 2 * np.pi), rng.uniform(0.8, 4.5)
                p = Path([x, y])
                p.push_line([x + r * np.cos(a0), y + r * np.sin(a0)])
                p.push_elliptical_arc([r, r], 0.0, a1 > np.pi, False, [x + r * np.cos(a0 + a1), y + r * np.sin(a0 + a1)])
            if cubic and k % 2 == 0:
                p.convert_quadratic_curves_to_cubic_curves()
            out.append(p)
        return out

    # role geometry in the group's local frame (about +-4 x +-3 model units)
    outlines = {
        0: Path.from_rounded_rect([0.0, 0.0], [4.0, 3.0], 0.8),          # clip level 1
        5: Path.from_ellipse([0.3, 0.1], [3.2, 2.3]),                     # clip level 2
        9: Path.from_circle([-0.4, 0.0], 2.4),                            # opacity group 0
        14: Path.from_rounded_rect([0.2, -0.1], [2.2, 1.6], 0.5),         # clip level 3
        18: Path.from_ellipse([0.4, 0.0], [1.8, 1.2]),                    # opacity group 1
    }
    paths: List[Path] = []
    for role in range(n_roles):
        main = outlines.get(role)
        if main is None:
            cx, cy = rng.uniform(-2.0, 2.0), rng.uniform(-1.5, 1.5)
            main = [Path.from_ellipse([cx, cy], [rng.uniform(0.6, 1.8), rng.uniform(0.5, 1.4)]),
                    Path.from_rounded_rect([cx, cy], [rng.uniform(0.8, 1.8), rng.uniform(0.6, 1.3)], 0.3),
                    Path.from_circle([cx, cy], rng.uniform(0.6, 1.5))][role % 3]
            if role % 4 == 1:
                main.convert_quadratic_curves_to_cubic_curves()
        paths.append(main)
        paths.extend(decorations(0.0, 0.0, 3.0, 2.2, paths_per_shape - 1, cubic=role % 2 == 0) if role not in outlines else
                     decorations(0.0, 0.0, 0.1, 0.1, paths_per_shape - 1, cubic=False))   # clip / group outlines: tiny inner decorations only
    soa = PathSoA.from_paths(paths)
    begin = np.arange(0, n_roles * paths_per_shape + 1, paths_per_shape, dtype=np.uint32)

    # per-copy placement: clip = ortho(W, H) * translate * rotate * scale, 2D affine in the instance mat4 (column vectors)
    centre = np.stack([rng.uniform(0, extent[0], n_instances), rng.uniform(0, extent[1], n_instances)], 1)
    scale = rng.uniform(instance_px[0], instance_px[1], n_instances) / 8.0   # the group is ~8 units wide
    angle = rng.uniform(0, 2 * np.pi, n_instances)
    c, s = np.cos(angle) * scale, np.sin(angle) * scale
    sx, sy = 2.0 / extent[0], -2.0 / extent[1]
    m = np.zeros((n_instances, 16))
    m[:, 0], m[:, 1] = sx * c, sy * s
    m[:, 4], m[:, 5] = -sx * s, sy * c
    m[:, 10] = 1.0
    m[:, 12], m[:, 13] = sx * centre[:, 0] - 1.0, sy * centre[:, 1] + 1.0
    m[:, 15] = 1.0
    transforms = np.repeat(m.astype(np.float32), n_roles, axis=0)
    colors = np.concatenate([rng.uniform(0, 1, (n_instances * n_roles, 3)), rng.uniform(0.35, 1.0, (n_instances * n_roles, 1))], 1).astype(np.float32)

    S, CLIP, UNCLIP, COLOR, SAVE, SCALE, RESTORE = range(7)

    def fills(lo, hi, depth, l0=0, l1=0):
        out = []
        for k in range(lo, hi):
            out += [(k, S, depth, l0, l1), (k, COLOR, depth, l0, l1)]
        return out

    group = ([(0, S, 0, 0, 0), (0, CLIP, 1, 0, 0)] + fills(1, 5, 1)
             + [(5, S, 1, 0, 0), (5, CLIP, 2, 0, 0)] + fills(6, 9, 2)
             + [(9, S, 2, 0, 0), (9, SAVE, 2, 0, 0), (9, SCALE, 2, 0, 0)] + fills(10, 14, 2)
             + [(14, S, 2, 0, 0), (14, CLIP, 3, 0, 0)] + fills(15, 18, 3)
             + [(18, S, 3, 1, 0), (18, SAVE, 3, 1, 0), (18, SCALE, 3, 1, 0)] + fills(19, 22, 3, 1, 0)
             + [(18, S, 3, 1, 1), (18, RESTORE, 3, 1, 1)]
             + [(14, S, 2, 1, 1), (14, UNCLIP, 2, 1, 1)]
             + [(9, S, 2, 1, 0), (9, RESTORE, 2, 1, 0)]
             + [(5, S, 1, 1, 0), (5, UNCLIP, 1, 1, 0)]
             + [(0, S, 0, 1, 0), (0, UNCLIP, 0, 1, 0)] + fills(22, 24, 0, 1, 0))
    g = np.array(group, np.int64)
    script = np.zeros((n_instances, len(g), 7), np.uint32)
    inst = (np.arange(n_instances)[:, None] * n_roles + g[None, :, 0])
    script[:, :, 0] = g[None, :, 0]
    script[:, :, 1] = inst
    script[:, :, 2] = inst + 1
    script[:, :, 3] = g[None, :, 1]
    script[:, :, 4:7] = g[None, :, 2:5]
    return ScriptedScene(soa, begin, [], extent[0], extent[1], np.ascontiguousarray(transforms), colors, script.reshape(-1, 7), "tiger_like", 2)


# ------------------------------------------------------------------------------------------------------ the reference's showcase
def showcase(extent: Tuple[int, int] = (640, 360), view_angles: Tuple[float, float] = (0.0, 0.0), view_distance: float = 5.0):
    """The scene of examples/showcase/main.rs, built with the host mirrors of the reference's own helpers: ONE Shape — the
    string "Hello World" through the text front-end (every glyph path reversed, main.rs:81-83) inside a rounded rectangle that
    is stroked with a dashed miter line (main.rs:58-92) — drawn 46 times: instance 0 with the camera's projection, a 5 x 9 grid
    of copies placed in 3D behind it (main.rs:163-201); the camera is `Translator(1, 0, 0, -view_distance / 2) * view_rotation`
    with `view_rotation = rotate_around_axis(a, y) * rotate_around_axis(b, x)` (main.rs:258-259). The demo renders it with
    4x MSAA, depth LessEqual + depth write, back-face culling (main.rs:28-52) and Stencil + Color per instance (main.rs:236-250).
    Returns (paths, shape_path_begin, dynamic_stroke_options, transforms [46, 16], colors [46, 4])."""
    import os
    from . import utils as U
    from .path import Cap, CurveApproximation, DashInterval, DynamicStrokeOptions, Join, Path, PathSoA, StrokeOptions
    from .text import Alignment, FixtureFace, Layout, Orientation, paths_of_text
    fixture = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "opensans_ascii.npz")
    face = FixtureFace(dict(np.load(fixture)))
    paths = paths_of_text(face, Layout(2.7, Orientation.LeftToRight, Alignment.Center, Alignment.Center), "Hello World")
    for p in paths:
        p.reverse()
    frame = Path.from_rounded_rect([0.0, 0.0], [5.8, 1.3], 0.5)
    frame.stroke_options = StrokeOptions(width=0.1, offset=0.0, miter_clip=1.0, closed=True, dynamic_stroke_options_group=0,
                                         curve_approximation=CurveApproximation.UniformTangentAngle(0.1))
    paths.insert(0, frame)
    dso = [DynamicStrokeOptions.Dashed(Join.Miter, [DashInterval(3.0, 4.0, Cap.Butt, Cap.Butt)], 0.0)]
    rotation = U.motor3d_product(U.rotor3d_to_motor3d(U.rotate_around_axis(view_angles[0], [0.0, 1.0, 0.0])),
                                 U.rotor3d_to_motor3d(U.rotate_around_axis(view_angles[1], [1.0, 0.0, 0.0])))
    camera = U.motor3d_product(U.translator3d(0.0, 0.0, -0.5 * view_distance), rotation)
    projection = U.matrix_multiplication(U.perspective_projection(np.pi * 0.5, extent[0] / extent[1], 1.0, 1000.0), U.motor3d_to_mat4(camera))
    rows, columns = 9, 5
    transforms, colors = [projection], [[1.0, 1.0, 1.0, 1.0]]
    for y in range(rows):
        for x in range(columns):
            place = U.translator3d((x + 0.5 - columns * 0.5) * 7.0, (y + 0.5 - rows * 0.5) * 3.0, -5.0)
            transforms.append(U.matrix_multiplication(projection, U.motor3d_to_mat4(place)))
            red, green = x / columns, y / rows
            colors.append([red, green, 1.0 - red - green, 1.0])
    soa = PathSoA.from_paths(paths)
    return (soa, np.array([0, soa.n_paths], np.uint32), dso, np.asarray(transforms, np.float32).reshape(-1, 16), np.asarray(colors, np.float32))
