"""Mirror of the reference's `forge/textsdf` package (font.go): a TrueType font turned into a polygon SDF tree,
backed by the C++ host layer (gsdf_b200/csrc/host/textsdf.cpp).

    f = textsdf.Font()
    f.Configure(RelativeGlyphTolerance=0.001)          # font.go:40
    f.LoadTTFBytes(open("iso-3098.ttf", "rb").read())  # font.go:54 (the reference embeds this font, embed.go:10-16)
    shape = f.TextLine(bld, "Abc123~")                  # font.go:89 -> Shader2D

The Builder is an explicit argument (the reference keeps it in FontConfig.Builder, font.go:25).
"""
import ctypes as C

import numpy as np

from ._lib import lib
from .gsdf import Shader


class FontError(ValueError):
    pass


class Font:
    def __init__(self):
        self._h = lib.gsdfh_font_new()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.gsdfh_font_free(h)

    def _err(self):
        return lib.gsdfh_font_err(self._h).decode("utf-8", "replace")

    def Configure(self, RelativeGlyphTolerance=0.0):
        if lib.gsdfh_font_configure(self._h, float(RelativeGlyphTolerance)) != 0:
            raise FontError(self._err())

    def LoadTTFBytes(self, ttf):
        ttf = bytes(ttf)
        if lib.gsdfh_font_load_ttf(self._h, ttf, len(ttf)) != 0:
            raise FontError(self._err())

    def TextLine(self, bld, s):
        nid = lib.gsdfh_font_textline(self._h, bld._h, s.encode("utf-8"))
        if nid < 0:
            raise FontError(self._err())
        return Shader(bld, nid)

    def Glyph(self, bld, c):
        nid = lib.gsdfh_font_glyph(self._h, bld._h, ord(c))
        if nid < 0:
            raise FontError(self._err())
        return Shader(bld, nid)

    def Kern(self, c0, c1):
        return float(lib.gsdfh_font_kern(self._h, ord(c0), ord(c1)))

    def AdvanceWidth(self, c):
        return float(lib.gsdfh_font_advance_width(self._h, ord(c)))

    # ---- parser introspection (tests)
    def scaleout(self):
        return float(lib.gsdfh_font_scaleout(self._h))

    def GlyphIndex(self, c):
        return int(lib.gsdfh_font_glyph_index(self._h, ord(c)))

    def Info(self):
        out = (C.c_int32 * 6)()
        if lib.gsdfh_font_info(self._h, out) != 0:
            raise FontError("no font loaded")
        return dict(unitsPerEm=out[0], numGlyphs=out[1], bounds=(out[2], out[3], out[4], out[5]))

    def GlyphSegments(self, glyph_index):
        """sfnt.LoadGlyph at ppem = unitsPerEm: int32 rows {op, x0,y0, x1,y1, x2,y2} (Y down)."""
        n = lib.gsdfh_font_glyph_segments(self._h, int(glyph_index), None, 0)
        if n < 0:
            raise FontError(self._err())
        out = np.zeros((max(n, 1), 7), np.int32)
        lib.gsdfh_font_glyph_segments(self._h, int(glyph_index), out.ctypes.data_as(C.POINTER(C.c_int32)), n)
        return out[:n]
