"""Text front-end (src/text.rs restated in contrast_renderer_b200/text.py): TrueType reader against the known facts of the
reference's demo font, the committed outline fixture, layout arithmetic and the vectorised path builder."""
import os

import numpy as np
import pytest

from contrast_renderer_b200 import _abi, scenes
from contrast_renderer_b200.path import PathSoA
from contrast_renderer_b200.text import (Alignment, Face, FixtureFace, Layout, Orientation, _emit_contour, calculate_aligned_positions, paths_of_glyph,
                                         paths_of_text, text_to_soa)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "opensans_ascii.npz")
FONT = "/root/reference/examples/fonts/OpenSans-Regular.ttf"
needs_font = pytest.mark.skipif(not os.path.exists(FONT), reason="the reference checkout (font file) is not on this machine")


@pytest.fixture(scope="module")
def fixture_face():
    return FixtureFace(dict(np.load(FIXTURE)))


@needs_font
def test_truetype_reader_against_known_font_facts():
    """SURVEY §2 row 19 (probed independently during the survey): 938 glyphs, unitsPerEm 2048, hhea 2189 / -600 / 0, xHeight
    1096, no kern table; printable ASCII averages 1.44 contours and 20.9 outline points per glyph; face.height() = 2789."""
    face = Face(open(FONT, "rb").read())
    assert (face.number_of_glyphs, face.units_per_em) == (938, 2048)
    assert (face.ascender(), face.descender(), face.line_gap(), face.x_height(), face.height()) == (2189, -600, 0, 1096, 2789)
    assert not face.has_kerning() and face.glyph_index("\ufffd") == 592
    contours = [face.glyph_points(face.glyph_index(chr(c))) for c in range(33, 127)]
    assert abs(sum(len(c) for c in contours) / 94 - 1.44) < 0.01
    assert abs(sum(len(p) for c in contours for p in c) / 94 - 20.9) < 0.01
    assert face.glyph_points(face.glyph_index(" ")) is None and face.glyph_hor_advance(face.glyph_index(" ")) > 0


@needs_font
def test_fixture_is_what_the_font_says(fixture_face):
    """tests/golden/opensans_ascii.npz (made by tests/golden/make_font_fixture.py) reproduces the font for every character it
    covers: glyph ids, advances, metrics and outlines."""
    face = Face(open(FONT, "rb").read())
    for c in list(range(0x20, 0x7F)) + [0xFFFD]:
        g = face.glyph_index(chr(c))
        assert fixture_face.glyph_index(chr(c)) == g and fixture_face.glyph_hor_advance(g) == face.glyph_hor_advance(g)
        a, b = face.outline_glyph(g), fixture_face.outline_glyph(g)
        assert (a is None) == (b is None)
        if a is not None:
            assert len(a.contours) == len(b.contours)
            for (s0, segs0), (s1, segs1) in zip(a.contours, b.contours):
                assert tuple(s0) == tuple(s1) and [(k, tuple(cp)) for k, cp in segs0] == [(k, tuple(cp)) for k, cp in segs1]
    assert (fixture_face.height(), fixture_face.line_gap(), fixture_face.descender(), fixture_face.x_height()) == (2789, 0, -600, 1096)


def test_outline_emission_rules():
    """ttf-parser glyf builder semantics, one case per rule."""
    on, off = True, False
    # all on-curve: lines, closed by an explicit line back to the start
    start, segs = _emit_contour([(0, 0, on), (10, 0, on), (10, 10, on)])
    assert start == (0, 0) and segs == [(0, (10, 0)), (0, (10, 10)), (0, (0, 0))]
    # on, off, on: one quadratic; the contour closes with a line
    start, segs = _emit_contour([(0, 0, on), (5, 8, off), (10, 0, on)])
    assert segs == [(1, (5, 8, 10, 0)), (0, (0, 0))]
    # two consecutive off-curve points: implied on-curve midpoint
    start, segs = _emit_contour([(0, 0, on), (0, 10, off), (10, 10, off), (10, 0, on)])
    assert segs == [(1, (0, 10, 5, 10)), (1, (10, 10, 10, 0)), (0, (0, 0))]
    # trailing off-curve point: the closing segment is a quadratic through it
    start, segs = _emit_contour([(0, 0, on), (10, 0, on), (5, 9, off)])
    assert segs == [(0, (10, 0)), (1, (5, 9, 0, 0))]
    # contour that starts off-curve: remembered, closed through it
    start, segs = _emit_contour([(5, 9, off), (0, 0, on), (10, 0, on)])
    assert start == (0, 0) and segs == [(0, (10, 0)), (1, (5, 9, 0, 0))]
    # contour made of off-curve points only: starts at the midpoint of the first two
    start, segs = _emit_contour([(0, 0, off), (10, 0, off), (10, 10, off), (0, 10, off)])
    assert start == (5, 0) and segs[0] == (1, (10, 0, 10, 5)) and segs[-1][0] == 1 and segs[-1][1][2:] == (5, 0)


def test_glyph_paths_are_closed_contours(fixture_face):
    """src/text.rs:61-104: one Path per contour, lines and integral quadratics only, ending where it started."""
    for ch in "OB%@g":
        paths = paths_of_glyph(fixture_face, fixture_face.glyph_index(ch))
        assert len(paths) == len(fixture_face.glyph_points(fixture_face.glyph_index(ch)))
        for p in paths:
            assert set(int(t) for t in p.segment_types) <= {_abi.CR_SEG_LINE, _abi.CR_SEG_INTEGRAL_QUADRATIC}
            assert np.array_equal(p.get_end(), p.start)
    assert len(paths_of_glyph(fixture_face, fixture_face.glyph_index("O"))) == 2
    assert paths_of_glyph(fixture_face, fixture_face.glyph_index(" ")) == []


def test_layout_arithmetic(fixture_face):
    """calculate_aligned_positions! (src/text.rs:145-230) by hand for OpenSans: i64 font units, truncating division."""
    adv = {c: fixture_face.glyph_hor_advance(fixture_face.glyph_index(c)) for c in "AVi"}
    lay = Layout(1.0, Orientation.LeftToRight, Alignment.Begin, Alignment.Baseline)
    extent, offset, lines = calculate_aligned_positions(fixture_face, lay, "AV\ni")
    width = adv["A"] + adv["V"]
    assert extent == [width, 2 * 2789] and len(lines) == 2 and lines[0][0] == 3 and lines[1][0] == 5
    # Begin: offset = -extent/2 on the major axis; Baseline: 0 on the minor axis minus (extent_minor - line_height)/2; y flipped
    x0 = -(width // 2)
    y_shift = (2 * 2789 - 2789) // 2
    assert [p for p, _ in lines[0][1]] == [[x0, y_shift], [x0 + adv["A"], y_shift], [x0 + width, y_shift]]
    assert [p for p, _ in lines[1][1]] == [[x0, -(2789 - y_shift)], [x0 + adv["i"], -(2789 - y_shift)]]
    assert [g for _, g in lines[1][1]] == [fixture_face.glyph_index("i"), 0]     # end-of-line marker carries glyph 0
    # the second tuple element is the OUTER `offset` (the loop's `let mut offset = offset;` shadows it, src/text.rs:216-229):
    # major component 0, minor component the alignment term alone (Baseline: 0), without the centring shift
    assert offset == [0, 0]
    # Center / Center: each line centred on its own extent, minor offset x_height / 2
    lay = Layout(1.0, Orientation.LeftToRight, Alignment.Center, Alignment.Center)
    _, offset, lines = calculate_aligned_positions(fixture_face, lay, "AV")
    assert offset == [0, -(1096 // 2)]   # sign_y = -1 (src/text.rs:150-158) times x_height / 2
    assert lines[0][1][0][0] == [-(width // 2) if width % 2 == 0 else -((width - 1) // 2), -(1096 // 2)]
    # RightToLeft mirrors x
    lay = Layout(1.0, Orientation.RightToLeft, Alignment.Begin, Alignment.Baseline)
    _, _, lines = calculate_aligned_positions(fixture_face, lay, "AV")
    assert lines[0][1][0][0][0] == (width // 2) and lines[0][1][1][0][0] == (width // 2) - adv["A"]


def test_paths_of_text_scale_and_vectorised_builder(fixture_face):
    """scale = size / 2789; text_to_soa is bit-identical to PathSoA.from_paths(paths_of_text(..))."""
    lay = Layout(2.7, Orientation.LeftToRight, Alignment.Center, Alignment.Center)   # the demo's layout (examples/showcase/main.rs:71-80)
    text = "Hello World"
    paths = paths_of_text(fixture_face, lay, text)
    assert len(paths) == 14                                             # 10 outlined glyphs, 14 contours
    soa, first = text_to_soa(fixture_face, lay, text)
    ref = PathSoA.from_paths(paths)
    for a, b in zip(soa.arrays(), ref.arrays()):
        assert np.array_equal(a, b)
    assert first.tolist() == [0, 1, 3, 4, 5, 7, 7, 8, 10, 11, 12]       # the space owns no path
    # H is 'scale' times its outline, translated: its stem height is cap height * 2.7 / 2789
    h = fixture_face.glyph_points(fixture_face.glyph_index("H"))[0]
    cap = max(p[1] for p in h) - min(p[1] for p in h)
    ys = np.concatenate([[paths[0].start[1]], *[s[1::2] for s in paths[0].line_segments]])
    assert abs((ys.max() - ys.min()) - cap * 2.7 / 2789) < 1e-5
    clip = np.array([[-100.0, -100.0], [-90.0, -100.0], [-90.0, -90.0], [-100.0, -90.0]])
    assert paths_of_text(fixture_face, lay, text, clip) == []            # everything outside the clipping polygon is discarded


def test_text_scene_is_deterministic_and_sized_like_config3():
    a, b = scenes.text_glyphs(2000, extent=(1024, 256), chars_per_line=160), scenes.text_glyphs(2000, extent=(1024, 256), chars_per_line=160)
    assert all(np.array_equal(x, y) for x, y in zip(a.paths.arrays(), b.paths.arrays()))
    assert a.n_shapes == 13 and abs(a.paths.n_paths / 2000 - 1.44) < 0.05
    assert set(np.unique(a.paths.segment_types)) == {_abi.CR_SEG_LINE, _abi.CR_SEG_INTEGRAL_QUADRATIC}
