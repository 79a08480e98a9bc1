"""Text front-end: font faces -> `Path`s, mirroring src/text.rs of the reference.

The reference delegates font parsing to the un-vendored crate `ttf-parser 0.14.0` (Cargo.lock:1475): `Face::from_slice`,
`glyph_index`, `glyph_hor_advance`, `outline_glyph` + `OutlineBuilder`, `height`, `line_gap`, `descender`, `x_height`,
`tables().kern`. `Face` below restates the part of it this path uses for TrueType (`glyf`) fonts — cmap formats 4 / 12,
simple and composite glyphs, the outline emission order of ttf-parser's glyf builder, format-0
kerning — and `FixtureFace` serves the same interface from a small table extracted from a font (tests/golden/), so that
benchmarks on machines without the font file lay out real glyph outlines.

`paths_of_glyph`, `Layout`, `paths_of_text` follow src/text.rs:97-104,134-143,236-263 including the i64 font-unit
layout arithmetic of `calculate_aligned_positions!` (:145-230) and the f32 operation order of `Path::transform`
(src/path.rs:387-439). `text_to_soa` is the vectorised equivalent of `PathSoA.from_paths(paths_of_text(..))` with an
outline cache per glyph (the reference re-outlines every occurrence).
"""
from __future__ import annotations

import enum
import struct
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi
from .path import Path, PathSoA


# ------------------------------------------------------------------------------------------------ outlines
@dataclass
class GlyphOutline:
    """A glyph as ttf-parser's OutlineBuilder sees it: per contour a start point and (kind, control points) segments in font
    units (f32). kind 0 = line_to (x, y), 1 = quad_to (x1, y1, x, y)."""
    contours: List[Tuple[Tuple[float, float], List[Tuple[int, Tuple[float, ...]]]]]


def _emit_contour(points: Sequence[Tuple[float, float, bool]]):
    """ttf-parser 0.14 glyf builder: on->on = line_to, off->on = quad_to(off, on), off->off =
    quad_to(off, mid), a contour starting off-curve starts at the midpoint of its first two off-curve points or closes
    through the remembered first off-curve point; every contour ends with an explicit closing segment."""
    f32 = np.float32
    first_on = first_off = last_off = None
    start = None
    segs: List[Tuple[int, Tuple[float, ...]]] = []

    def mid(a, b):
        return (f32(a[0] + (b[0] - a[0]) * f32(0.5)), f32(a[1] + (b[1] - a[1]) * f32(0.5)))

    for x, y, on in points:
        p = (f32(x), f32(y))
        if first_on is None:
            if on:
                first_on = p
                start = p
            elif first_off is not None:
                m = mid(first_off, p)
                first_on = m
                last_off = p
                start = m
            else:
                first_off = p
        elif last_off is not None and on:
            segs.append((1, (*last_off, *p)))
            last_off = None
        elif last_off is not None:
            m = mid(last_off, p)
            segs.append((1, (*last_off, *m)))
            last_off = p
        elif on:
            segs.append((0, p))
        else:
            last_off = p
    # finish_contour
    if first_off is not None and last_off is not None:
        m = mid(last_off, first_off)
        segs.append((1, (*last_off, *m)))
        last_off = None
    if first_on is not None and first_off is not None:
        segs.append((1, (*first_off, *first_on)))
    elif first_on is not None and last_off is not None:
        segs.append((1, (*last_off, *first_on)))
    elif first_on is not None:
        segs.append((0, first_on))
    if start is None:
        return None
    return (start, segs)


class Face:
    """The subset of `ttf_parser::Face` that src/text.rs uses, for TrueType outlines."""

    def __init__(self, data: bytes):
        self.data = data
        n_tables = struct.unpack_from(">H", data, 4)[0]
        self.tables: Dict[str, Tuple[int, int]] = {}
        for i in range(n_tables):
            tag, _, off, length = struct.unpack_from(">4sIII", data, 12 + 16 * i)
            self.tables[tag.decode("latin1")] = (off, length)
        head = self.tables["head"][0]
        self.units_per_em = struct.unpack_from(">H", data, head + 18)[0]
        self.index_to_loc_format = struct.unpack_from(">h", data, head + 50)[0]
        self.number_of_glyphs = struct.unpack_from(">H", data, self.tables["maxp"][0] + 4)[0]
        hhea = self.tables["hhea"][0]
        self.hhea_ascender, self.hhea_descender, self.hhea_line_gap = struct.unpack_from(">hhh", data, hhea + 4)
        self.number_of_h_metrics = struct.unpack_from(">H", data, hhea + 34)[0]
        self._os2 = self.tables.get("OS/2")
        self._cmap = self._parse_cmap()
        self._kern = self._parse_kern()
        self._outline_cache: Dict[int, Optional[GlyphOutline]] = {}

    # ---- metrics (ttf-parser: hhea values unless OS/2 USE_TYPO_METRICS is set)
    def _use_typo(self) -> bool:
        if not self._os2:
            return False
        return bool(struct.unpack_from(">H", self.data, self._os2[0] + 62)[0] & (1 << 7))

    def ascender(self) -> int:
        return struct.unpack_from(">h", self.data, self._os2[0] + 68)[0] if self._use_typo() else self.hhea_ascender

    def descender(self) -> int:
        return struct.unpack_from(">h", self.data, self._os2[0] + 70)[0] if self._use_typo() else self.hhea_descender

    def line_gap(self) -> int:
        return struct.unpack_from(">h", self.data, self._os2[0] + 72)[0] if self._use_typo() else self.hhea_line_gap

    def height(self) -> int:
        return self.ascender() - self.descender()

    def x_height(self) -> Optional[int]:
        if not self._os2 or struct.unpack_from(">H", self.data, self._os2[0])[0] < 2:
            return None
        return struct.unpack_from(">h", self.data, self._os2[0] + 86)[0]

    def vertical_height(self) -> Optional[int]:
        return None   # no vhea support (OpenSans has none; src/text.rs:160 falls back to 0)

    def vertical_line_gap(self) -> Optional[int]:
        return None

    # ---- cmap
    def _parse_cmap(self):
        off = self.tables["cmap"][0]
        n = struct.unpack_from(">H", self.data, off + 2)[0]
        best = None
        for i in range(n):
            pid, eid, sub = struct.unpack_from(">HHI", self.data, off + 4 + 8 * i)
            fmt = struct.unpack_from(">H", self.data, off + sub)[0]
            if (pid == 3 and eid in (1, 10)) or pid == 0:
                if fmt == 12 or (fmt == 4 and best is None):
                    best = (fmt, off + sub)
        return best

    def glyph_index(self, char: str) -> Optional[int]:
        code = ord(char)
        if self._cmap is None:
            return None
        fmt, off = self._cmap
        d = self.data
        if fmt == 4:
            if code > 0xFFFF:
                return None
            seg_x2 = struct.unpack_from(">H", d, off + 6)[0]
            ends = off + 14
            starts = ends + seg_x2 + 2
            deltas = starts + seg_x2
            ranges = deltas + seg_x2
            for i in range(seg_x2 // 2):
                end = struct.unpack_from(">H", d, ends + 2 * i)[0]
                if code <= end:
                    start = struct.unpack_from(">H", d, starts + 2 * i)[0]
                    if code < start:
                        return None
                    delta = struct.unpack_from(">h", d, deltas + 2 * i)[0]
                    ro = struct.unpack_from(">H", d, ranges + 2 * i)[0]
                    if ro == 0:
                        gid = (code + delta) & 0xFFFF
                    else:
                        gid = struct.unpack_from(">H", d, ranges + 2 * i + ro + 2 * (code - start))[0]
                        if gid != 0:
                            gid = (gid + delta) & 0xFFFF
                    return gid or None
            return None
        n_groups = struct.unpack_from(">I", d, off + 12)[0]
        for i in range(n_groups):
            s, e, g = struct.unpack_from(">III", d, off + 16 + 12 * i)
            if s <= code <= e:
                return g + (code - s)
        return None

    # ---- hmtx / kern
    def glyph_hor_advance(self, gid: int) -> Optional[int]:
        off = self.tables["hmtx"][0]
        i = min(gid, self.number_of_h_metrics - 1)
        return struct.unpack_from(">H", self.data, off + 4 * i)[0]

    def glyph_ver_advance(self, gid: int) -> Optional[int]:
        return None

    def _parse_kern(self):
        if "kern" not in self.tables:
            return None
        off = self.tables["kern"][0]
        version, n = struct.unpack_from(">HH", self.data, off)
        if version != 0 or n == 0:
            return None
        _, length, coverage = struct.unpack_from(">HHH", self.data, off + 4)
        if (coverage >> 8) != 0:
            return None
        n_pairs = struct.unpack_from(">H", self.data, off + 10)[0]
        pairs = {}
        for i in range(n_pairs):
            l, r, v = struct.unpack_from(">HHh", self.data, off + 18 + 6 * i)
            pairs[(l, r)] = v
        return pairs

    def glyphs_kerning(self, left: int, right: int) -> Optional[int]:
        return None if self._kern is None else self._kern.get((left, right))

    def has_kerning(self) -> bool:
        return self._kern is not None

    # ---- glyf
    def _glyph_range(self, gid: int) -> Tuple[int, int]:
        loca = self.tables["loca"][0]
        if self.index_to_loc_format == 0:
            a, b = struct.unpack_from(">HH", self.data, loca + 2 * gid)
            return 2 * a, 2 * b
        return struct.unpack_from(">II", self.data, loca + 4 * gid)

    def glyph_points(self, gid: int, depth: int = 0) -> Optional[List[List[Tuple[float, float, bool]]]]:
        """Contours of a glyph as lists of (x, y, on_curve) in font units; composites are flattened (transformed points)."""
        if gid >= self.number_of_glyphs or depth > 32:
            return None
        a, b = self._glyph_range(gid)
        if a == b:
            return None
        d, g = self.data, self.tables["glyf"][0] + a
        n_contours = struct.unpack_from(">h", d, g)[0]
        if n_contours >= 0:
            ends = struct.unpack_from(f">{n_contours}H", d, g + 10)
            n_points = ends[-1] + 1 if n_contours else 0
            p = g + 10 + 2 * n_contours
            p += 2 + struct.unpack_from(">H", d, p)[0]
            flags: List[int] = []
            while len(flags) < n_points:
                f = d[p]
                p += 1
                flags.append(f)
                if f & 8:
                    r = d[p]
                    p += 1
                    flags.extend([f] * r)
            xs, ys = [], []
            v = 0
            for f in flags:
                if f & 2:
                    dx = d[p]
                    p += 1
                    v += dx if f & 16 else -dx
                elif not f & 16:
                    v += struct.unpack_from(">h", d, p)[0]
                    p += 2
                xs.append(v)
            v = 0
            for f in flags:
                if f & 4:
                    dy = d[p]
                    p += 1
                    v += dy if f & 32 else -dy
                elif not f & 32:
                    v += struct.unpack_from(">h", d, p)[0]
                    p += 2
                ys.append(v)
            contours, s = [], 0
            for e in ends:
                contours.append([(float(xs[i]), float(ys[i]), bool(flags[i] & 1)) for i in range(s, e + 1)])
                s = e + 1
            return contours
        # composite glyph
        p = g + 10
        out: List[List[Tuple[float, float, bool]]] = []
        while True:
            flags, sub = struct.unpack_from(">HH", d, p)
            p += 4
            if flags & 1:
                e, f_ = struct.unpack_from(">hh", d, p)
                p += 4
            else:
                e, f_ = struct.unpack_from(">bb", d, p)
                p += 2
            m = [1.0, 0.0, 0.0, 1.0]
            if flags & 8:
                m[0] = m[3] = struct.unpack_from(">h", d, p)[0] / 16384.0
                p += 2
            elif flags & 0x40:
                sx, sy = struct.unpack_from(">hh", d, p)
                m[0], m[3] = sx / 16384.0, sy / 16384.0
                p += 4
            elif flags & 0x80:
                m = [v / 16384.0 for v in struct.unpack_from(">hhhh", d, p)]
                p += 8
            tx, ty = (float(e), float(f_)) if flags & 2 else (0.0, 0.0)
            sub_contours = self.glyph_points(sub, depth + 1)
            for c in sub_contours or []:
                out.append([(float(np.float32(m[0] * x + m[2] * y + tx)), float(np.float32(m[1] * x + m[3] * y + ty)), on) for x, y, on in c])
            if not flags & 0x20:
                break
        return out or None

    def outline_glyph(self, gid: int) -> Optional[GlyphOutline]:
        if gid not in self._outline_cache:
            contours = self.glyph_points(gid)
            if contours is None:
                self._outline_cache[gid] = None
            else:
                emitted = [_emit_contour(c) for c in contours if len(c) >= 2]
                self._outline_cache[gid] = GlyphOutline([c for c in emitted if c is not None])
        return self._outline_cache[gid]


class FixtureFace(Face):
    """A `Face` served from a table extracted from a font (tests/golden/make_font_fixture.py): raw glyf points of a set of
    characters plus the metrics src/text.rs reads. No font file needed."""

    def __init__(self, fixture: Dict[str, np.ndarray]):   # noqa: super().__init__ deliberately not called: there is no sfnt
        meta = fixture["metrics"]
        (self.units_per_em, asc, desc, gap, xh, self.number_of_glyphs) = [int(v) for v in meta]
        self.hhea_ascender, self.hhea_descender, self.hhea_line_gap = asc, desc, gap
        self._x_height = xh
        self._codes = {int(c): int(g) for c, g in zip(fixture["codepoints"], fixture["glyph_ids"])}
        self._advance = {int(g): int(a) for g, a in zip(fixture["glyph_ids"], fixture["advances"])}
        self._points: Dict[int, List[List[Tuple[float, float, bool]]]] = {}
        pts, ends, begin = fixture["points"], fixture["contour_ends"], fixture["glyph_contour_begin"]
        for k, g in enumerate(fixture["glyph_ids"]):
            contours, s = [], int(fixture["glyph_point_begin"][k])
            for e in ends[int(begin[k]):int(begin[k + 1])]:
                contours.append([(float(x), float(y), bool(on)) for x, y, on in pts[s:int(e)]])
                s = int(e)
            self._points[int(g)] = contours
        self._kern = None
        self._outline_cache = {}

    def _use_typo(self) -> bool:
        return False

    def x_height(self) -> Optional[int]:
        return self._x_height

    def glyph_index(self, char: str) -> Optional[int]:
        return self._codes.get(ord(char))

    def glyph_hor_advance(self, gid: int) -> Optional[int]:
        return self._advance.get(gid)

    def glyph_points(self, gid: int, depth: int = 0):
        c = self._points.get(gid)
        return c if c else None


def paths_of_glyph(face: Face, glyph_id: int) -> List[Path]:
    """src/text.rs:97-104: one Path per contour, each ending with its explicit closing segment."""
    outline = face.outline_glyph(glyph_id)
    paths: List[Path] = []
    for start, segs in (outline.contours if outline else []):
        path = Path(start)
        for kind, cp in segs:
            if kind == 0:
                path.push_line(cp)
            else:
                path.push_integral_quadratic_curve([cp[0:2], cp[2:4]])
        paths.append(path)
    return paths


# -------------------------------------------------------------------------------------------------- layout
class Orientation(enum.Enum):   # src/text.rs:106-116
    RightToLeft = 0
    LeftToRight = 1
    TopToBottom = 2
    BottomToTop = 3


class Alignment(enum.Enum):     # src/text.rs:119-129
    Begin = 0
    Baseline = 1
    Center = 2
    End = 3


@dataclass
class Layout:                   # src/text.rs:132-143
    size: float
    orientation: Orientation = Orientation.LeftToRight
    major_alignment: Alignment = Alignment.Begin
    minor_alignment: Alignment = Alignment.Baseline


def _div_trunc(a: int, b: int) -> int:
    """Rust i64 division truncates toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def calculate_aligned_positions(face: Face, layout: Layout, text: str):
    """calculate_aligned_positions! (src/text.rs:145-230): positions in integer font units. Returns (extent, offset, lines)
    with lines = [(line_range_end, [((x, y), glyph_id), ...])]; the last entry of every line is the end-of-line marker."""
    replacement = face.glyph_index("�")
    major_axis, sign_x, sign_y = {Orientation.RightToLeft: (0, -1, -1), Orientation.LeftToRight: (0, 1, -1), Orientation.TopToBottom: (1, 1, -1),
                                  Orientation.BottomToTop: (1, 1, 1)}[layout.orientation]
    if major_axis == 0:
        line_minor_extent, line_gap = face.height(), face.line_gap()
    else:
        line_minor_extent, line_gap = face.vertical_height() or 0, face.vertical_line_gap() or 0
    lines = []
    line_major_extent = 0
    extent = [0, 0]
    glyph_positions = []
    prev = None
    index = 0
    for ch in text:
        index += 1
        pos = list(extent)
        pos[major_axis] = line_major_extent
        if ch == "\n":
            glyph_positions.append((pos, 0))
            lines.append((index, glyph_positions))
            glyph_positions = []
            extent[major_axis] = max(extent[major_axis], line_major_extent)
            extent[1 - major_axis] += line_minor_extent + line_gap
            line_major_extent = 0
            prev = None
        else:
            gid = face.glyph_index(ch)
            if gid is None:
                gid = replacement
            if gid is None:
                raise ValueError(f"no glyph for {ch!r} and no U+FFFD replacement glyph")   # `.unwrap()` panics in the reference
            if face.has_kerning() and prev is not None:
                k = face.glyphs_kerning(prev, gid)
                if k is not None:
                    line_major_extent += k
            prev = gid
            adv = face.glyph_hor_advance(gid) if major_axis == 0 else face.glyph_ver_advance(gid)
            if adv is not None:
                line_major_extent += adv
            glyph_positions.append((pos, gid))
    pos = list(extent)
    pos[major_axis] = line_major_extent
    glyph_positions.append((pos, 0))
    lines.append((index + 1, glyph_positions))
    extent[major_axis] = max(extent[major_axis], line_major_extent)
    extent[1 - major_axis] += line_minor_extent
    offset = [0, 0]
    if layout.minor_alignment == Alignment.Begin:
        offset[1 - major_axis] = -face.descender()
    elif layout.minor_alignment == Alignment.Baseline:
        offset[1 - major_axis] = 0
    elif layout.minor_alignment == Alignment.Center:
        offset[1 - major_axis] = _div_trunc(face.x_height(), 2)
    else:
        offset[1 - major_axis] = -line_minor_extent
    for _, positions in lines:
        lme = positions[-1][0][major_axis]
        off = list(offset)
        if layout.major_alignment == Alignment.Begin:
            off[major_axis] = _div_trunc(-extent[major_axis], 2)
        elif layout.major_alignment in (Alignment.Baseline, Alignment.Center):
            off[major_axis] = _div_trunc(-lme, 2)
        else:
            off[major_axis] = _div_trunc(extent[major_axis], 2) - lme
        off[1 - major_axis] -= _div_trunc(extent[1 - major_axis] - line_minor_extent, 2)
        for p, _ in positions:
            p[0] = sign_x * (p[0] + off[0])
            p[1] = sign_y * (p[1] + off[1])
    # the macro's loop shadows `offset` (`let mut offset = offset;`, src/text.rs:216): the tuple returns the OUTER one
    return extent, [sign_x * offset[0], sign_y * offset[1]], lines


def _sat_overlap(a: np.ndarray, b: np.ndarray) -> bool:
    """do_convex_polygons_overlap (src/utils.rs:85-99): separating axis test on two convex polygons."""
    for poly, other in ((a, b), (b, a)):
        for i in range(len(poly)):
            e = poly[(i + 1) % len(poly)] - poly[i]
            n = np.array([e[1], -e[0]])
            pa, pb = poly @ n, other @ n
            if pa.max() < pb.min() or pb.max() < pa.min():
                return False
    return True


def glyph_occurrences(face: Face, layout: Layout, text: str):
    """[(x, y, glyph_id)] in font units for every character of `text` in order (end-of-line markers dropped)."""
    _, _, lines = calculate_aligned_positions(face, layout, text)
    return [(p[0], p[1], gid) for _, positions in lines for p, gid in positions[:-1]]


def paths_of_text(face: Face, layout: Layout, text: str, clipping_area: Optional[np.ndarray] = None) -> List[Path]:
    """src/text.rs:236-263: every glyph occurrence is outlined, scaled by size / face.height() and translated."""
    scale = np.float32(np.float32(layout.size) / np.float32(face.height()))
    result: List[Path] = []
    for x, y, gid in glyph_occurrences(face, layout, text):
        if clipping_area is not None:
            pts = [q for c in (face.glyph_points(gid) or []) for q in c]
            if pts:
                xs, ys = [q[0] for q in pts], [q[1] for q in pts]
                x0, y0, x1, y1 = [np.float32(np.float32(v) * scale) for v in (min(xs) + x, min(ys) + y, max(xs) + x, max(ys) + y)]
                if not _sat_overlap(np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], np.float64), np.asarray(clipping_area, np.float64)):
                    continue
        tx, ty = np.float32(np.float32(x) * scale), np.float32(np.float32(y) * scale)
        for path in paths_of_glyph(face, gid):
            _transform_path(path, scale, tx, ty)
            result.append(path)
    return result


def _transform_path(path: Path, scale: np.float32, tx: np.float32, ty: np.float32) -> None:
    """Path::transform(scale, &translate2d([tx, ty])) (src/path.rs:387-439): p' = t + p * scale in f32, in that order."""
    def tp(v: np.ndarray) -> np.ndarray:
        out = v.astype(np.float32).copy()
        out[0::2] = tx + out[0::2] * scale
        out[1::2] = ty + out[1::2] * scale
        return out + np.float32(0.0)
    path.start = tp(path.start)
    path.line_segments = [tp(s) for s in path.line_segments]
    path.integral_quadratic_curve_segments = [tp(s) for s in path.integral_quadratic_curve_segments]
    path.integral_cubic_curve_segments = [tp(s) for s in path.integral_cubic_curve_segments]


def text_to_soa(face: Face, layout: Layout, text: str) -> Tuple[PathSoA, np.ndarray]:
    """Vectorised `PathSoA.from_paths(paths_of_text(face, layout, text, None))`: identical arrays, but each distinct glyph is
    outlined once. Also returns, per glyph occurrence, the index of its first path (occurrences without an outline, e.g.
    spaces, own no paths), so that callers can group glyphs into Shapes."""
    occ = glyph_occurrences(face, layout, text)
    scale = np.float32(np.float32(layout.size) / np.float32(face.height()))
    gids = np.array([o[2] for o in occ], np.int64)
    uniq = sorted(set(gids.tolist()))
    slot = {g: i for i, g in enumerate(uniq)}
    # per-glyph templates in font units
    t_paths, t_start, t_segcount, t_types = [], [], [], []
    t_rows: List[List[np.ndarray]] = [[], []]     # line rows [x, y], quad rows [x1, y1, x, y]
    t_type_count = [[], []]
    for g in uniq:
        outline = face.outline_glyph(g)
        contours = outline.contours if outline else []
        t_paths.append(len(contours))
        for start, segs in contours:
            t_start.append(start)
            t_segcount.append(len(segs))
            t_types.extend(k for k, _ in segs)
            t_rows[0].extend(cp for k, cp in segs if k == 0)
            t_rows[1].extend(cp for k, cp in segs if k == 1)
            t_type_count[0].append(sum(1 for k, _ in segs if k == 0))
            t_type_count[1].append(sum(1 for k, _ in segs if k == 1))
    t_paths = np.array(t_paths, np.int64)
    path_off = np.concatenate([[0], np.cumsum(t_paths)])                         # first template path of each glyph slot
    t_start = np.array(t_start, np.float32).reshape(-1, 2)
    t_segcount = np.array(t_segcount, np.int64)
    seg_off = np.concatenate([[0], np.cumsum(t_segcount)])
    t_types = np.array(t_types, np.uint8)
    t_lines = np.array(t_rows[0], np.float32).reshape(-1, 2)
    t_quads = np.array(t_rows[1], np.float32).reshape(-1, 4)
    cnt = [np.array(t_type_count[0], np.int64), np.array(t_type_count[1], np.int64)]
    row_off = [np.concatenate([[0], np.cumsum(c)]) for c in cnt]

    def expand(counts: np.ndarray, offsets: np.ndarray):
        """For items with `counts` elements starting at `offsets` in a template array: gather indices and the owning item."""
        total = int(counts.sum())
        owner = np.repeat(np.arange(len(counts)), counts)
        first = np.cumsum(counts) - counts
        return offsets[owner] + (np.arange(total) - first[owner]), owner

    o_slot = np.array([slot[g] for g in gids.tolist()], np.int64)
    o_paths = t_paths[o_slot]
    # paths of the text, in order: template path index + owning occurrence
    p_tmpl, p_occ = expand(o_paths, path_off[o_slot])
    tx = (np.array([o[0] for o in occ], np.float32) * scale)[p_occ]
    ty = (np.array([o[1] for o in occ], np.float32) * scale)[p_occ]
    start = np.stack([tx + t_start[p_tmpl, 0] * scale, ty + t_start[p_tmpl, 1] * scale], 1).astype(np.float32) + np.float32(0.0)
    seg_counts = t_segcount[p_tmpl]
    s_tmpl, s_path = expand(seg_counts, seg_off[p_tmpl])
    seg_types = t_types[s_tmpl]
    payload = [np.zeros((0, w), np.float32) for w in _abi.SEGMENT_FLOATS]
    type_begin = np.zeros((5, len(p_tmpl) + 1), np.uint32)
    for t, tmpl_rows in ((0, t_lines), (1, t_quads)):
        counts = cnt[t][p_tmpl]
        r_tmpl, r_path = expand(counts, row_off[t][p_tmpl])
        rows = tmpl_rows[r_tmpl].copy()
        rows[:, 0::2] = tx[r_path, None] + rows[:, 0::2] * scale
        rows[:, 1::2] = ty[r_path, None] + rows[:, 1::2] * scale
        payload[t] = rows.astype(np.float32) + np.float32(0.0)
        np.cumsum(counts, out=type_begin[t, 1:])
    segment_begin = np.zeros(len(p_tmpl) + 1, np.uint32)
    np.cumsum(seg_counts, out=segment_begin[1:])
    soa = PathSoA(start, segment_begin, seg_types, type_begin, payload, np.zeros(len(p_tmpl), PathSoA.STROKE_DTYPE))
    occ_first_path = (np.cumsum(o_paths) - o_paths).astype(np.int64)
    return soa, occ_first_path
