"""Test helper: the ISO-3098 glyph subset fixture (tests/golden/iso3098_subset.json) as a loadable .ttf."""
import json
import os

import ttfwriter

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FONT = "/root/reference/forge/textsdf/iso-3098.ttf"


def fixture():
    with open(os.path.join(HERE, "golden", "iso3098_subset.json")) as fp:
        return json.load(fp)


def subset_ttf(compact=True, long_loca=False):
    fx = fixture()
    return ttfwriter.write_ttf(fx["glyphs"], fx["unitsPerEm"], fx["bbox"], compact=compact, long_loca=long_loca)


def text_scene(bld, text="Abc123~", tol=0.001):
    """examples/image-text/text.go:24-34 on the fixture font."""
    from gsdf_b200 import textsdf
    f = textsdf.Font()
    f.Configure(RelativeGlyphTolerance=tol)
    f.LoadTTFBytes(subset_ttf())
    return f.TextLine(bld, text)
