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
