"""Extracts tests/golden/opensans_ascii.npz from the reference's demo font (examples/fonts/OpenSans-Regular.ttf), so that the
text front-end and the config-3 benchmark scene can lay out real OpenSans outlines on machines where /root/reference does not
exist. Run in the build container: python tests/golden/make_font_fixture.py

Stored: for U+0020..U+007E and U+FFFD the glyph id, advance width, raw `glyf` points (x, y, on-curve) and contour ends, plus
unitsPerEm / hhea ascender, descender, lineGap / OS/2 sxHeight / numGlyphs. Nothing else of the font is copied."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from contrast_renderer_b200.text import Face  # noqa: E402

FONT = "/root/reference/examples/fonts/OpenSans-Regular.ttf"


def main():
    face = Face(open(FONT, "rb").read())
    codes = list(range(0x20, 0x7F)) + [0xFFFD]
    gids, adv, pts, ends, cbegin, pbegin = [], [], [], [], [0], []
    for c in codes:
        g = face.glyph_index(chr(c))
        assert g is not None, hex(c)
        gids.append(g)
        adv.append(face.glyph_hor_advance(g))
        pbegin.append(len(pts))
        for contour in face.glyph_points(g) or []:
            pts.extend((x, y, int(on)) for x, y, on in contour)
            ends.append(len(pts))
        cbegin.append(len(ends))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "opensans_ascii.npz")
    np.savez_compressed(out, codepoints=np.array(codes, np.uint32), glyph_ids=np.array(gids, np.uint32), advances=np.array(adv, np.uint32),
                        points=np.array(pts, np.int16).reshape(-1, 3), contour_ends=np.array(ends, np.uint32),
                        glyph_contour_begin=np.array(cbegin, np.uint32), glyph_point_begin=np.array(pbegin, np.uint32),
                        metrics=np.array([face.units_per_em, face.ascender(), face.descender(), face.line_gap(), face.x_height(), face.number_of_glyphs], np.int32))
    print(out, os.path.getsize(out), "bytes;", len(pts), "points,", len(ends), "contours")


if __name__ == "__main__":
    main()
