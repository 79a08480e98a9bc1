"""Host-side mirror of the reference's path model (src/path.rs:13-261) and its packing into the C-ABI's
structure-of-arrays (`cr_path_soa`, include/contrast_b200.h).

Names, field meaning and defaults follow the reference so that code written against `contrast_renderer::path`
reads the same here: `Path`, the five segment kinds, `StrokeOptions`, `DynamicStrokeOptions`, `Join`, `Cap`,
`DashInterval`, `CurveApproximation`.
"""
from __future__ import annotations

import ctypes as C
import enum
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Union

import numpy as np

from . import _abi


def safe_float(value) -> np.float32:
    """`SafeFloat::from` (src/safe_float.rs:44-52): assert finite, canonicalise -0.0 to +0.0."""
    v = np.float32(value)
    if not np.isfinite(v):
        raise ValueError("SafeFloat requires a finite value")
    return np.float32(0.0) if v == 0 else v


def safe_vec(values) -> np.ndarray:
    return np.array([safe_float(v) for v in values], dtype=np.float32)


class SegmentType(enum.IntEnum):  # src/path.rs:56-67
    Line = 0
    IntegralQuadraticCurve = 1
    IntegralCubicCurve = 2
    RationalQuadraticCurve = 3
    RationalCubicCurve = 4


class Join(enum.IntEnum):  # src/path.rs:71-82
    Miter = 0
    Bevel = 1
    Round = 2


class Cap(enum.IntEnum):  # src/path.rs:86-101
    Square = 0
    Round = 1
    Out = 2
    In = 3
    Right = 4
    Left = 5
    Butt = 6


@dataclass
class DashInterval:  # src/path.rs:105-118
    gap_start: float
    gap_end: float
    dash_start: Cap = Cap.Butt
    dash_end: Cap = Cap.Butt


@dataclass
class DynamicStrokeOptions:
    """`enum DynamicStrokeOptions` (src/path.rs:127-149). Use the `Dashed` / `Solid` constructors."""
    dashed: bool
    join: Join
    start: Cap = Cap.Butt
    end: Cap = Cap.Butt
    pattern: List[DashInterval] = field(default_factory=list)
    phase: float = 0.0

    @staticmethod
    def Dashed(join: Join, pattern: Sequence[DashInterval], phase: float) -> "DynamicStrokeOptions":
        return DynamicStrokeOptions(True, join, pattern=list(pattern), phase=phase)

    @staticmethod
    def Solid(join: Join, start: Cap, end: Cap) -> "DynamicStrokeOptions":
        return DynamicStrokeOptions(False, join, start=start, end=end)

    def to_c(self) -> _abi.DynamicStrokeOptionsC:
        c = _abi.DynamicStrokeOptionsC()
        c.dashed = 1 if self.dashed else 0
        c.join = int(self.join)
        c.start = int(self.start)
        c.end = int(self.end)
        c.pattern_len = len(self.pattern)
        c.phase = float(safe_float(self.phase))
        if len(self.pattern) > _abi.CR_DASH_PATTERN_CAPACITY:
            # cannot even be expressed in the C struct: same error the library reports for > MAX_DASH_INTERVALS
            c.pattern_len = _abi.CR_DASH_PATTERN_CAPACITY
        for i, d in enumerate(self.pattern[: _abi.CR_DASH_PATTERN_CAPACITY]):
            c.pattern[i].gap_start = float(safe_float(d.gap_start))
            c.pattern[i].gap_end = float(safe_float(d.gap_end))
            c.pattern[i].dash_start = int(d.dash_start)
            c.pattern[i].dash_end = int(d.dash_end)
        return c


def dynamic_stroke_options_array(options: Sequence[DynamicStrokeOptions]):
    arr = (_abi.DynamicStrokeOptionsC * max(1, len(options)))()
    for i, o in enumerate(options):
        arr[i] = o.to_c()
    return arr


@dataclass
class CurveApproximation:  # src/path.rs:153-167
    uniform_tangent_angle: Optional[float] = None
    uniformly_spaced_parameters: Optional[int] = None

    @staticmethod
    def UniformTangentAngle(a: float) -> "CurveApproximation":
        return CurveApproximation(uniform_tangent_angle=float(safe_float(a)))

    @staticmethod
    def UniformlySpacedParameters(n: int) -> "CurveApproximation":
        return CurveApproximation(uniformly_spaced_parameters=int(n))


@dataclass
class StrokeOptions:  # src/path.rs:171-201
    width: float
    offset: float = 0.0
    miter_clip: float = 1.0
    closed: bool = False
    dynamic_stroke_options_group: int = 0
    curve_approximation: CurveApproximation = field(default_factory=lambda: CurveApproximation.UniformTangentAngle(0.1))

    def legalize(self) -> None:
        """src/path.rs:196-200"""
        self.width = abs(self.width)
        self.offset = min(0.5, max(-0.5, self.offset))
        self.miter_clip = abs(self.miter_clip)


@dataclass
class LineSegment:
    control_points: Sequence[Sequence[float]]  # [1][2]


@dataclass
class IntegralQuadraticCurveSegment:
    control_points: Sequence[Sequence[float]]  # [2][2]


@dataclass
class IntegralCubicCurveSegment:
    control_points: Sequence[Sequence[float]]  # [3][2]


@dataclass
class RationalQuadraticCurveSegment:
    weight: float
    control_points: Sequence[Sequence[float]]  # [2][2]


@dataclass
class RationalCubicCurveSegment:
    weights: Sequence[float]  # [4], including the start
    control_points: Sequence[Sequence[float]]  # [3][2]


class Path:
    """`struct Path` (src/path.rs:213-230): per-type segment storage plus `segment_types` interleave order."""

    def __init__(self, start=(0.0, 0.0), stroke_options: Optional[StrokeOptions] = None):
        self.stroke_options = stroke_options
        self.start = safe_vec(start)
        self.line_segments: List[np.ndarray] = []
        self.integral_quadratic_curve_segments: List[np.ndarray] = []
        self.integral_cubic_curve_segments: List[np.ndarray] = []
        self.rational_quadratic_curve_segments: List[np.ndarray] = []
        self.rational_cubic_curve_segments: List[np.ndarray] = []
        self.segment_types: List[int] = []

    # src/path.rs:234-261
    def push_line(self, segment: Union[LineSegment, Sequence[float]]) -> None:
        pts = segment.control_points[0] if isinstance(segment, LineSegment) else segment
        self.line_segments.append(safe_vec(pts))
        self.segment_types.append(SegmentType.Line)

    def push_integral_quadratic_curve(self, segment: Union[IntegralQuadraticCurveSegment, Sequence[Sequence[float]]]) -> None:
        pts = segment.control_points if isinstance(segment, IntegralQuadraticCurveSegment) else segment
        self.integral_quadratic_curve_segments.append(safe_vec(np.asarray(pts, dtype=np.float64).reshape(-1)))
        self.segment_types.append(SegmentType.IntegralQuadraticCurve)

    def push_integral_cubic_curve(self, segment: Union[IntegralCubicCurveSegment, Sequence[Sequence[float]]]) -> None:
        pts = segment.control_points if isinstance(segment, IntegralCubicCurveSegment) else segment
        self.integral_cubic_curve_segments.append(safe_vec(np.asarray(pts, dtype=np.float64).reshape(-1)))
        self.segment_types.append(SegmentType.IntegralCubicCurve)

    def push_rational_quadratic_curve(self, segment: RationalQuadraticCurveSegment) -> None:
        flat = [segment.weight] + list(np.asarray(segment.control_points, dtype=np.float64).reshape(-1))
        self.rational_quadratic_curve_segments.append(safe_vec(flat))
        self.segment_types.append(SegmentType.RationalQuadraticCurve)

    def push_rational_cubic_curve(self, segment: RationalCubicCurveSegment) -> None:
        flat = list(segment.weights) + list(np.asarray(segment.control_points, dtype=np.float64).reshape(-1))
        self.rational_cubic_curve_segments.append(safe_vec(flat))
        self.segment_types.append(SegmentType.RationalCubicCurve)

    def get_end(self) -> np.ndarray:
        """src/path.rs:266-290"""
        if not self.segment_types:
            return self.start.copy()
        t = self.segment_types[-1]
        store = (self.line_segments, self.integral_quadratic_curve_segments, self.integral_cubic_curve_segments,
                 self.rational_quadratic_curve_segments, self.rational_cubic_curve_segments)[t]
        return store[-1][-2:].copy()

    def close(self) -> None:
        """src/path.rs:620-627: no-op if |end - start|^2 <= ERROR_MARGIN, else a line back to the start."""
        d = self.get_end() - self.start
        if float(np.float32(d[0] * d[0] + d[1] * d[1])) <= 1e-4:
            return
        self.push_line(self.start)

    def push_quarter_ellipse(self, tangent_crossing, to) -> None:
        """src/path.rs:630-635: a rational quadratic with middle weight 1/sqrt(2)."""
        self.push_rational_quadratic_curve(RationalQuadraticCurveSegment(np.float32(0.70710678118654752440), [tangent_crossing, to]))

    # ---- src/path.rs:376-617: append, transform, reverse, conversions --------------------------------------------------
    _N_WEIGHTS = (0, 0, 0, 1, 4)   # leading weight floats of a stored segment, by SegmentType

    def _stores(self):
        return (self.line_segments, self.integral_quadratic_curve_segments, self.integral_cubic_curve_segments,
                self.rational_quadratic_curve_segments, self.rational_cubic_curve_segments)

    def append(self, other: "Path") -> None:
        """src/path.rs:376-384: concatenates the segment vectors and leaves `other` empty. The reference forgets
        `segment_types`, which makes the appended segments unreachable; here they are appended
        too, the evident intent."""
        for mine, theirs in zip(self._stores(), other._stores()):
            mine.extend(theirs)
            theirs.clear()
        self.segment_types.extend(other.segment_types)
        other.segment_types.clear()

    def transform(self, scale: float, motor) -> None:
        """src/path.rs:387-439: every control point (start included) through `motor2d_to_mat3(motor)` whose two diagonal
        entries are multiplied by `scale` (so `scale` is a uniform scale only for rotation-free motors, as in the reference)."""
        from . import utils
        t = utils.motor2d_to_mat3(motor)
        t[0][0] *= np.float32(scale)
        t[1][1] *= np.float32(scale)

        def transform_point(q):
            return safe_vec([t[2][0] + q[0] * t[0][0] + q[1] * t[1][0], t[2][1] + q[0] * t[0][1] + q[1] * t[1][1]])

        self.start = transform_point(self.start)
        for kind, store in enumerate(self._stores()):
            w = Path._N_WEIGHTS[kind]
            for seg in store:
                for i in range(w, len(seg), 2):
                    seg[i:i + 2] = transform_point(seg[i:i + 2])

    def reverse(self) -> None:
        """src/path.rs:445-488: swaps start and end, reverses every segment and the segment order."""
        previous = self.start.copy()
        cursors = [0] * 5
        for kind in self.segment_types:
            seg = self._stores()[kind][cursors[kind]]
            cursors[kind] += 1
            w = Path._N_WEIGHTS[kind]
            if kind == SegmentType.RationalCubicCurve:
                seg[0:4] = seg[0:4][::-1].copy()
            if kind in (SegmentType.IntegralCubicCurve, SegmentType.RationalCubicCurve):   # control_points.swap(0, 1)
                a = seg[w:w + 2].copy()
                seg[w:w + 2] = seg[w + 2:w + 4]
                seg[w + 2:w + 4] = a
            end = seg[-2:].copy()
            seg[-2:] = previous
            previous = end
        self.start = previous
        self.segment_types.reverse()
        for store in self._stores():
            store.reverse()

    def convert_integral_curves_to_rational_curves(self) -> None:
        """src/path.rs:492-531: integral quadratics / cubics become rational ones with unit weights, in path order."""
        cursors = [0] * 5
        rq: List[np.ndarray] = []
        rc: List[np.ndarray] = []
        for i, kind in enumerate(self.segment_types):
            seg = self._stores()[kind][cursors[kind]]
            cursors[kind] += 1
            if kind == SegmentType.IntegralQuadraticCurve:
                rq.append(np.concatenate([np.ones(1, np.float32), seg]))
                self.segment_types[i] = SegmentType.RationalQuadraticCurve
            elif kind == SegmentType.IntegralCubicCurve:
                rc.append(np.concatenate([np.ones(4, np.float32), seg]))
                self.segment_types[i] = SegmentType.RationalCubicCurve
            elif kind == SegmentType.RationalQuadraticCurve:
                rq.append(seg)
            elif kind == SegmentType.RationalCubicCurve:
                rc.append(seg)
        self.rational_quadratic_curve_segments[:] = rq
        self.rational_cubic_curve_segments[:] = rc
        self.integral_quadratic_curve_segments.clear()
        self.integral_cubic_curve_segments.clear()

    def convert_quadratic_curves_to_cubic_curves(self) -> None:
        """src/path.rs:535-617: degree elevation; the rational case elevates the homogeneous control points and
        normalises the two inner ones (weights [1, w1, w2, 1])."""
        f = np.float32
        cursors = [0] * 5
        ic: List[np.ndarray] = []
        rc: List[np.ndarray] = []
        previous = self.start.copy()
        two_thirds = f(2.0) / f(3.0)
        for i, kind in enumerate(self.segment_types):
            seg = self._stores()[kind][cursors[kind]]
            cursors[kind] += 1
            if kind == SegmentType.IntegralQuadraticCurve:
                a, b = seg[0:2], seg[2:4]
                # reference: previous + (a - previous) * 2.0 / 3.0 (two roundings: multiply, then divide)
                c0 = previous + (a - previous) * f(2.0) / f(3.0)
                c1 = b + (a - b) * f(2.0) / f(3.0)
                ic.append(safe_vec(np.concatenate([c0, c1, b])))
                self.segment_types[i] = SegmentType.IntegralCubicCurve
            elif kind == SegmentType.IntegralCubicCurve:
                ic.append(seg)
            elif kind == SegmentType.RationalQuadraticCurve:
                w = seg[0]
                p0 = np.array([1.0, previous[0], previous[1]], np.float32)
                p1 = np.array([w, seg[1] * w, seg[2] * w], np.float32)
                p2 = np.array([1.0, seg[3], seg[4]], np.float32)
                n0 = p0 + (p1 - p0) * two_thirds
                n1 = p2 + (p1 - p2) * two_thirds
                rc.append(safe_vec([1.0, n0[0], n1[0], 1.0, n0[1] / n0[0], n0[2] / n0[0], n1[1] / n1[0], n1[2] / n1[0], seg[3], seg[4]]))
                self.segment_types[i] = SegmentType.RationalCubicCurve
            elif kind == SegmentType.RationalCubicCurve:
                rc.append(seg)
            previous = seg[-2:].copy()
        self.integral_cubic_curve_segments[:] = ic
        self.rational_cubic_curve_segments[:] = rc
        self.integral_quadratic_curve_segments.clear()
        self.rational_quadratic_curve_segments.clear()

    # ---- src/path.rs:637-815: arcs and shape constructors ---------------------------------------------------------------
    def push_elliptical_arc(self, half_extent, rotation: float, large_arc: bool, sweep: bool, to) -> None:
        """src/path.rs:638-703: SVG "arc to" (https://www.w3.org/TR/SVG/implnote.html) as rational quadratics of at most
        120 degrees each."""
        from . import utils as U
        f = np.float32
        radii = np.array([0.0, abs(f(half_extent[0])), abs(f(half_extent[1]))], np.float32)
        if radii[1] == 0.0 or radii[2] == 0.0:
            self.push_line(to)
            return
        frm = U.vec_to_point(self.get_end())
        to_p = U.vec_to_point(safe_vec(to))
        rotor = U.rotate2d(rotation)
        vertex_unoriented = (to_p - frm) * f(0.5)                       # .dual(): component-wise identity
        vertex = U.motor_transform_plane(U.motor_inverse(rotor), vertex_unoriented)
        vertex_squared = vertex * vertex
        radii_squared = radii * radii
        scale_factor_squared = vertex_squared[1] / radii_squared[1] + vertex_squared[2] / radii_squared[2]
        if scale_factor_squared > 1.0:                                  # radii too small to span from -> to: scale them up
            radii = radii * np.sqrt(scale_factor_squared)
            radii_squared = radii * radii
        one_over_radii = np.array([0.0, f(1.0) / radii[1], f(1.0) / radii[2]], np.float32)
        rsvs = radii_squared[1] * vertex_squared[2] + radii_squared[2] * vertex_squared[1]
        offset = np.sqrt(max(f(0.0), (radii_squared[1] * radii_squared[2] - rsvs) / rsvs))
        if large_arc == sweep:
            offset = -offset
        center_offset_unoriented = radii * U.rotate_90_degree_clockwise(vertex * one_over_radii) * offset
        center = (to_p + frm) * f(0.5) + U.motor_transform_plane(rotor, center_offset_unoriented)
        start_normal = (-vertex - center_offset_unoriented) * one_over_radii
        end_normal = (vertex - center_offset_unoriented) * one_over_radii
        polar_start = U.complex_signum([start_normal[1], start_normal[2]])
        polar_end = U.complex_signum([end_normal[1], end_normal[2]])
        polar_range = U.complex_div(polar_end, polar_start)
        small_arc = U.complex_arg(polar_range)
        if small_arc < 0.0:
            polar_range = np.array([polar_range[0], -polar_range[1]], np.float32)   # reversal
            small_arc = -small_arc
        angle = small_arc
        if large_arc:
            angle = angle - U.TAU
        step_radians = f(math.pi) * f(2.0) / f(3.0)
        steps = int(math.ceil(float(abs(angle) / step_radians)))
        if large_arc != sweep:
            angle = -angle
        polar_step = U.complex_powf(polar_range, angle / (small_arc * f(steps)))
        half_polar_step_back = U.complex_powf(polar_step, -0.5)
        weight = np.cos(abs(angle) / f(steps) * f(0.5))
        tangent_crossing_radii = radii * (f(1.0) / weight)
        for i in range(1, steps + 1):
            interpolated = U.complex_mul(polar_start, U.complex_powi(polar_step, i))
            v_unoriented = np.array([0.0, interpolated[0], interpolated[1]], np.float32) * radii
            v = center + U.motor_transform_plane(rotor, v_unoriented)
            interpolated = U.complex_mul(interpolated, half_polar_step_back)
            tc_unoriented = np.array([0.0, interpolated[0], interpolated[1]], np.float32) * tangent_crossing_radii
            tc = center + U.motor_transform_plane(rotor, tc_unoriented)
            self.push_rational_quadratic_curve(RationalQuadraticCurveSegment(weight, [U.point_to_vec(tc), U.point_to_vec(v)]))

    @staticmethod
    def from_polygon(vertices, stroke_options: Optional[StrokeOptions] = None) -> "Path":
        """src/path.rs:706-718"""
        result = Path(vertices[0], stroke_options)
        for v in vertices[1:]:
            result.push_line(v)
        return result

    @staticmethod
    def from_regular_polygon(center, radius: float, rotation: float, vertex_count: int, stroke_options: Optional[StrokeOptions] = None) -> "Path":
        """src/path.rs:721-728"""
        f = np.float32
        vertices = []
        for i in range(vertex_count):
            angle = f(rotation) + f(i) / f(vertex_count) * f(math.pi) * f(2.0)
            vertices.append([f(center[0]) + f(radius) * np.cos(angle), f(center[1]) + f(radius) * np.sin(angle)])
        return Path.from_polygon(vertices, stroke_options)

    @staticmethod
    def from_rect(center, half_extent, stroke_options: Optional[StrokeOptions] = None) -> "Path":
        """src/path.rs:731-738"""
        f = np.float32
        cx, cy, hx, hy = f(center[0]), f(center[1]), f(half_extent[0]), f(half_extent[1])
        return Path.from_polygon([[cx - hx, cy - hy], [cx - hx, cy + hy], [cx + hx, cy + hy], [cx + hx, cy - hy]], stroke_options)

    @staticmethod
    def from_rounded_rect(center, half_extent, radius: float, stroke_options: Optional[StrokeOptions] = None) -> "Path":
        """src/path.rs:741-776: four lines and four quarter circles, starting at the end of the last rounding."""
        f = np.float32
        cx, cy, hx, hy, r = f(center[0]), f(center[1]), f(half_extent[0]), f(half_extent[1]), f(radius)
        vertices = [
            ([cx - hx + r, cy - hy], [cx - hx, cy - hy], [cx - hx, cy - hy + r]),
            ([cx - hx, cy + hy - r], [cx - hx, cy + hy], [cx - hx + r, cy + hy]),
            ([cx + hx - r, cy + hy], [cx + hx, cy + hy], [cx + hx, cy + hy - r]),
            ([cx + hx, cy - hy + r], [cx + hx, cy - hy], [cx + hx - r, cy - hy]),
        ]
        result = Path(vertices[3][2], stroke_options)
        for frm, corner, to in vertices:
            result.push_line(frm)
            result.push_quarter_ellipse(corner, to)
        return result

    @staticmethod
    def from_ellipse(center, half_extent, stroke_options: Optional[StrokeOptions] = None) -> "Path":
        """src/path.rs:779-808: four quarter ellipses."""
        f = np.float32
        cx, cy, hx, hy = f(center[0]), f(center[1]), f(half_extent[0]), f(half_extent[1])
        vertices = [
            ([cx - hx, cy - hy], [cx - hx, cy]),
            ([cx - hx, cy + hy], [cx, cy + hy]),
            ([cx + hx, cy + hy], [cx + hx, cy]),
            ([cx + hx, cy - hy], [cx, cy - hy]),
        ]
        result = Path(vertices[3][1], stroke_options)
        for corner, to in vertices:
            result.push_quarter_ellipse(corner, to)
        return result

    @staticmethod
    def from_circle(center, radius: float, stroke_options: Optional[StrokeOptions] = None) -> "Path":
        """src/path.rs:811-813"""
        return Path.from_ellipse(center, [radius, radius], stroke_options)


@dataclass
class PathSoA:
    """Numpy structure-of-arrays for a set of paths; `as_c()` yields the `cr_path_soa` the C-ABI consumes."""
    start: np.ndarray            # [n, 2] f32
    segment_begin: np.ndarray    # [n + 1] u32
    segment_types: np.ndarray    # [n_segments] u8
    type_begin: np.ndarray       # [5, n + 1] u32
    segments: List[np.ndarray]   # five arrays [n_t, SEGMENT_FLOATS[t]] f32
    stroke_options: np.ndarray   # [n] structured (24 B), see STROKE_DTYPE

    STROKE_DTYPE = np.dtype([("width", "<f4"), ("offset", "<f4"), ("miter_clip", "<f4"), ("flags", "<u4"), ("group", "<u4"),
                             ("approx", "<u4")])

    @property
    def n_paths(self) -> int:
        return int(self.start.shape[0])

    @property
    def n_segments(self) -> int:
        return int(self.segment_types.shape[0])

    def input_bytes(self) -> int:
        """B_in of SURVEY §8d: 8 + 24*[stroked] per path, 1 + sizeof(segment) per segment."""
        stroked = int(((self.stroke_options["flags"] & _abi.CR_STROKE_FLAG_STROKED) != 0).sum())
        seg = sum(int(a.shape[0]) * (1 + 4 * w) for a, w in zip(self.segments, _abi.SEGMENT_FLOATS))
        return 8 * self.n_paths + 24 * stroked + seg

    def any_stroked(self) -> bool:
        return bool(((self.stroke_options["flags"] & _abi.CR_STROKE_FLAG_STROKED) != 0).any())

    def arrays(self):
        return [self.start, self.segment_begin, self.segment_types, self.type_begin, *self.segments, self.stroke_options]

    def as_c(self, memory_space: int = _abi.CR_MEM_HOST, pointers=None) -> _abi.PathSoAC:
        """`pointers`: optional list of 10 raw addresses (e.g. device pointers) in `arrays()` order."""
        c = _abi.PathSoAC()
        c.n_paths = self.n_paths
        c.n_segments = self.n_segments
        c.memory_space = memory_space
        if pointers is None:
            for a in self.arrays():
                assert a.flags["C_CONTIGUOUS"]
            pointers = [a.ctypes.data for a in self.arrays()]
            if not self.any_stroked():
                pointers[9] = None   # `stroke_options: None` for every Path: nothing to send
        (c.start, c.segment_begin, c.segment_types, c.type_begin, c.line_segments, c.integral_quadratic, c.integral_cubic,
         c.rational_quadratic, c.rational_cubic, c.stroke_options) = pointers
        return c

    @staticmethod
    def stroke_record(so: Optional[StrokeOptions]):
        if so is None:
            return (0.0, 0.0, 0.0, 0, 0, 0)
        flags = _abi.CR_STROKE_FLAG_STROKED | (_abi.CR_STROKE_FLAG_CLOSED if so.closed else 0)
        ca = so.curve_approximation
        if ca.uniform_tangent_angle is not None:
            flags |= _abi.CR_STROKE_FLAG_UNIFORM_TANGENT_ANGLE
            approx = int(np.float32(ca.uniform_tangent_angle).view(np.uint32))
        else:
            approx = int(ca.uniformly_spaced_parameters)
        return (float(safe_float(so.width)), float(safe_float(so.offset)), float(safe_float(so.miter_clip)), flags,
                int(so.dynamic_stroke_options_group), approx)

    @staticmethod
    def from_paths(paths: Sequence[Path]) -> "PathSoA":
        n = len(paths)
        start = np.zeros((n, 2), np.float32)
        segment_begin = np.zeros(n + 1, np.uint32)
        type_begin = np.zeros((5, n + 1), np.uint32)
        seg_types: List[int] = []
        stores: List[List[np.ndarray]] = [[], [], [], [], []]
        stroke = np.zeros(n, PathSoA.STROKE_DTYPE)
        for i, p in enumerate(paths):
            start[i] = p.start
            seg_types.extend(int(t) for t in p.segment_types)
            segment_begin[i + 1] = len(seg_types)
            per_type = (p.line_segments, p.integral_quadratic_curve_segments, p.integral_cubic_curve_segments,
                        p.rational_quadratic_curve_segments, p.rational_cubic_curve_segments)
            for t in range(5):
                stores[t].extend(per_type[t])
                type_begin[t, i + 1] = len(stores[t])
            stroke[i] = PathSoA.stroke_record(p.stroke_options)
        segments = [np.ascontiguousarray(np.array(stores[t], dtype=np.float32).reshape(-1, _abi.SEGMENT_FLOATS[t])) for t in range(5)]
        return PathSoA(start, segment_begin, np.array(seg_types, dtype=np.uint8), type_begin, segments, stroke)


def quarter_circle_weight() -> float:
    """Weight of the middle control point of a rational quadratic quarter circle (src/path.rs:633)."""
    return 1.0 / math.sqrt(2.0)
