"""forge/textsdf mirror (gsdf_b200/textsdf.py -> csrc/host/textsdf.cpp): font parsing, glyph outlines, TextLine.

The reference's own test (forge/textsdf/glyph_test.go:13-31) only checks that "Abp8" builds and renders; here the
outline decoding is additionally pinned against FreeType rasterisations of the same glyphs (tests/golden/
iso3098_raster.npz, made by tests/golden/make_golden_font.py) and against an independent glyf reader.
"""
import os

import numpy as np
import pytest

import fontfix
import ttfwriter
from gsdf_b200 import gsdf, textsdf


def _expected_segments(contours):
    """Restates the contour walk on the fixture's raw points (Y up) -> list of (op, args...) with Y down."""
    segs = []
    for c in contours:
        pts = [((x, -y), bool(on)) for x, y, on in c]
        mid = lambda a, b: (int((a[0] + b[0]) / 2), int((a[1] + b[1]) / 2))  # noqa: E731  truncation toward zero
        first_on = first_off = last_off = None
        for p, on in pts:
            if first_on is None:
                if on:
                    first_on = p
                    segs.append((0, p))
                elif first_off is None:
                    first_off = p
                else:
                    first_on = mid(first_off, p)
                    last_off = p
                    segs.append((0, first_on))
            elif last_off is None:
                if on:
                    segs.append((1, p))
                else:
                    last_off = p
            else:
                if on:
                    segs.append((2, last_off, p))
                    last_off = None
                else:
                    segs.append((2, last_off, mid(last_off, p)))
                    last_off = p
        if first_off is not None and last_off is not None:
            segs.append((2, last_off, mid(last_off, first_off)))
            last_off = None
        if first_off is None and last_off is None:
            segs.append((1, first_on))
        elif first_off is None:
            segs.append((2, last_off, first_on))
        else:
            segs.append((2, first_off, first_on))
    return segs


@pytest.mark.parametrize("compact,long_loca", [(True, False), (False, True)])
def test_parser_reads_back_the_fixture_font(compact, long_loca):
    fx = fontfix.fixture()
    f = textsdf.Font()
    f.LoadTTFBytes(fontfix.subset_ttf(compact=compact, long_loca=long_loca))
    info = f.Info()
    assert info["unitsPerEm"] == fx["unitsPerEm"] == 1000
    xmin, ymin, xmax, ymax = fx["bbox"]
    assert info["bounds"] == (xmin, -ymax, xmax, -ymin)  # sfnt.Font.Bounds flips Y
    assert f.scaleout() == np.float32(1) / np.float32(min(xmax - xmin, ymax - ymin))  # font.go:208-212 -> 1/933
    for ch, g in fx["glyphs"].items():
        gi = f.GlyphIndex(ch)
        assert gi > 0
        got = f.GlyphSegments(gi)
        want = _expected_segments(g["contours"])
        assert len(got) == len(want)
        for row, w in zip(got, want):
            assert row[0] == w[0]
            flat = [v for p in w[1:] for v in p]
            assert list(row[1:1 + len(flat)]) == flat
        assert f.AdvanceWidth(ch) == np.float32(g["advance"]) * np.float32(f.scaleout())
    assert f.GlyphIndex("Z") == 0  # unmapped rune -> .notdef, like sfnt.GlyphIndex
    assert f.Kern("A", "b") == 0.0  # no kern table: sfnt.Kern errors and TextLine ignores it (font.go:125)


@pytest.mark.skipif(not os.path.exists(fontfix.REF_FONT), reason="reference font not on this machine")
def test_parser_on_the_reference_font_matches_the_fixture():
    f = textsdf.Font()
    f.LoadTTFBytes(open(fontfix.REF_FONT, "rb").read())
    g = textsdf.Font()
    g.LoadTTFBytes(fontfix.subset_ttf())
    assert f.Info()["unitsPerEm"] == g.Info()["unitsPerEm"] and f.Info()["bounds"] == g.Info()["bounds"]
    assert f.Info()["numGlyphs"] == 203
    for ch in fontfix.fixture()["glyphs"]:
        assert np.array_equal(f.GlyphSegments(f.GlyphIndex(ch)), g.GlyphSegments(g.GlyphIndex(ch)))
        assert f.AdvanceWidth(ch) == g.AdvanceWidth(ch)
    # every glyph of the real font decodes (203 glyphs, simple outlines only)
    for gi in range(203):
        f.GlyphSegments(gi)


def test_glyph_sdf_sign_agrees_with_freetype(oracle):
    """Sign of the polygon SDF (oracle evaluation of the tree built by Font.Glyph) vs FreeType's monochrome raster."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "iso3098_raster.npz"))
    W, H, X0, YB, ppem, S = [int(v) for v in z["meta"]]
    f = textsdf.Font()
    f.Configure(RelativeGlyphTolerance=0.001)
    f.LoadTTFBytes(fontfix.subset_ttf())
    so = f.scaleout()
    ii, jj = np.meshgrid(np.arange(W), np.arange(H))
    # sampled pixel centre -> font units -> glyph coordinates (font units * scaleout, Y up)
    ux = (S * ii + S // 2 + 0.5 - X0) * (1000.0 / ppem)
    uy = (YB - (S * jj + S // 2 + 0.5)) * (1000.0 / ppem)
    pos = np.stack([ux * so, uy * so], -1).astype(np.float32).reshape(-1, 2)
    unit = np.float32(so)  # one font unit
    for ch in fontfix.fixture()["glyphs"]:
        bld = gsdf.Builder()
        s = f.Glyph(bld, ch)
        d = oracle.Tree.from_shader(s).eval2(pos).reshape(H, W)
        ink = np.unpackbits(z["u%04x" % ord(ch)])[:W * H].reshape(H, W).astype(bool)
        assert ink.sum() > 200, ch
        # chord tolerance 0.001 = 0.93 font units, FreeType's mono hinting moves outlines by up to ~3 units at this
        # size ('~' cap): everything farther than 4 units (0.4 % of the em, strokes are ~70 wide) must agree
        far = np.abs(d) > 4 * unit
        assert np.array_equal((d < 0)[far], ink[far]), "glyph %r: %d sign mismatches" % (ch, int(((d < 0) != ink)[far].sum()))
        assert ((d < 0) != ink).mean() < 0.001


def test_textline_structure_and_bounds(bld):
    s = fontfix.text_scene(bld)
    fx = fontfix.fixture()
    nodes = bld.tree_nodes()
    root = nodes[s.id]
    assert root.kind == gsdf.K["UNION2D"] and root.nchild == 7  # one translate2D per glyph, nested unions flattened (operations2d.go:27-35)
    so = np.float32(1) / np.float32(933)
    x = 0
    kids = bld.tree_children()
    for k, ch in enumerate("Abc123~"):
        t = nodes[kids[root.child_off + k]]
        assert t.kind == gsdf.K["TRANSLATE2D"]
        assert t.fparam[0] == np.float32(np.float32(x) * so) and t.fparam[1] == 0.0  # font.go:132
        x += fx["glyphs"][ch]["advance"]
    mn, mx = s.Bounds()
    assert -0.01 < mn[0] < 0.05 and mx[0] <= x * so and 0.5 < mx[1] - mn[1] < 1.2
    # 'A' has a hole (2 contours -> Difference2D, font.go:250-254); 'c' is one polygon
    a = nodes[kids[nodes[kids[root.child_off + 0]].child_off]]
    assert a.kind == gsdf.K["DIFF2D"]
    c = nodes[kids[nodes[kids[root.child_off + 2]].child_off]]
    assert c.kind == gsdf.K["POLY2D"]


def test_textline_errors(bld):
    f = textsdf.Font()
    with pytest.raises(textsdf.FontError, match="no font loaded"):
        f.TextLine(bld, "A")
    with pytest.raises(textsdf.FontError, match="invalid RelativeGlyphTolerance"):
        f.Configure(RelativeGlyphTolerance=1.0)  # font.go:41-43
    with pytest.raises(textsdf.FontError):
        f.LoadTTFBytes(b"not a font")
    f.LoadTTFBytes(fontfix.subset_ttf())
    with pytest.raises(textsdf.FontError, match="not graphic"):
        f.TextLine(bld, "A\tb")  # unicode.IsGraphic('\t') is false (font.go:96-98)
    with pytest.raises(textsdf.FontError, match="no text provided"):
        f.TextLine(bld, "   ")  # font.go:136-139
    with pytest.raises(textsdf.FontError, match="glyph has no contours"):
        f.TextLine(bld, "AZ")  # unmapped rune -> empty .notdef glyph (font.go:236-238)
    one = f.TextLine(bld, "A")  # a single glyph is returned as is (font.go:134-135)
    assert bld.tree_nodes()[one.id].kind == gsdf.K["TRANSLATE2D"]
    # spaces advance the pen without adding a shape
    two = f.TextLine(bld, "A b")
    assert bld.tree_nodes()[two.id].nchild == 2


def test_tolerance_controls_the_sampling(bld):
    def nverts(tol):
        b = gsdf.Builder()
        f = textsdf.Font()
        f.Configure(RelativeGlyphTolerance=tol)
        f.LoadTTFBytes(fontfix.subset_ttf())
        s = f.Glyph(b, "8")
        return sum(n.aux_cnt // 2 for n in b.tree_nodes() if n.kind == gsdf.K["POLY2D"]), s
    coarse, _ = nverts(0.15)   # default tolerance: chords between on-curve points only
    fine, _ = nverts(0.001)    # examples/image-text/text.go:27
    finer, _ = nverts(0.00001)
    assert coarse < fine <= finer
    # SampleBisect(poly, 4) appends at most 2^4-1 interior points per curve segment (font.go:311)
    fx = fontfix.fixture()
    nseg = sum(len(c) for c in fx["glyphs"]["8"]["contours"])
    assert finer <= 16 * nseg


def test_text_image_matches_the_oracle_golden(oracle, bld):
    """Config 5 at a reduced size through the oracle: frozen bits of the 2-D evaluator path on the text scene."""
    s = fontfix.text_scene(bld)
    mn, mx = s.Bounds()
    img = oracle.Tree.from_shader(s).image_eval2(mn, mx, 96, 24)
    assert np.isfinite(img).all() and (img < 0).any() and (img > 0).any()
    inside = (img < 0).mean()
    assert 0.05 < inside < 0.5
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "text_image.npz"))
    assert np.array_equal(np.concatenate([mn, mx]).view(np.uint32), z["bounds"].view(np.uint32))
    assert np.array_equal(img.view(np.uint32), z["dist"].view(np.uint32))


def _ops(flat):
    import struct
    words = np.frombuffer(flat["blob"], np.uint32, offset=32).reshape(-1, 4)
    out, pc = [], 0
    while pc < len(words):
        op, ln = int(words[pc, 0]) & 0xff, (int(words[pc, 0]) >> 8) & 0xff
        out.append((pc, op, ln, words[pc]))
        pc += ln
    return out


def test_box_guards_are_emitted_for_the_text_union(bld, monkeypatch):
    """include/gsdf_program.h "box guards": the text union gets one CULL_UB2D, one BBOX_GUARD2D(MIN) per glyph and one
    BBOX_GUARD2D(DIFF) per hole; every guard's target is its combiner; GSDF_NO_GUARDS=1 removes them all."""
    OP_MIN, OP_DIFF, OP_CULL, OP_BBOX = 20, 22, 51, 52  # enum gsdf_opcode (include/gsdf_program.h)
    s = fontfix.text_scene(bld)
    ops = _ops(bld.flatten(s))
    kinds = [o[1] for o in ops]
    assert kinds.count(OP_CULL) == 1 and kinds.index(OP_CULL) == 0
    guards = [o for o in ops if o[1] == OP_BBOX]
    starts = {o[0]: o[1] for o in ops}
    nmin = sum(1 for g in guards if int(g[3][1]) & 0xff == 2)
    ndiff = sum(1 for g in guards if int(g[3][1]) & 0xff == 1)
    assert nmin == 7          # one per glyph of "Abc123~"
    assert ndiff == 2         # 'A' and 'b' have a hole (Difference2D, font.go:250-254)
    for g in guards:
        kind, target = int(g[3][1]) & 0xff, int(g[3][1]) >> 8
        assert target > g[0] and starts[target] == (OP_MIN if kind == 2 else OP_DIFF)
    assert kinds.count(OP_MIN) == 7  # U is an extra operand of the fold: one MIN per glyph
    monkeypatch.setenv("GSDF_NO_GUARDS", "1")
    plain = [o[1] for o in _ops(bld.flatten(s))]
    assert OP_CULL not in plain and OP_BBOX not in plain and plain.count(OP_MIN) == 6


def test_damaged_fonts_fail_cleanly():
    """LoadTTFBytes takes caller-supplied bytes (font.go:54): truncated or corrupted fonts must end in an error or a
    parsed font, never in a crash. 600 seeded mutations of the fixture font (byte flips, truncations, 4-byte splats);
    fuzzing the same mutations under AddressSanitizer found an out-of-bounds read in the cmap lookup, fixed by
    bounds-checking every multi-byte read (host/textsdf.h)."""
    base = bytearray(fontfix.subset_ttf())
    rng = np.random.default_rng(3)
    parsed = failed = 0
    for it in range(600):
        b = bytearray(base)
        k = int(rng.integers(0, 3))
        if k == 0:
            for _ in range(int(rng.integers(1, 8))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif k == 1:
            b = b[:int(rng.integers(0, len(b)))]
        else:
            i = int(rng.integers(0, len(b) - 4))
            b[i:i + 4] = bytes(rng.integers(0, 256, 4, dtype=np.uint8))
        f = textsdf.Font()
        f.Configure(RelativeGlyphTolerance=0.01)
        try:
            f.LoadTTFBytes(bytes(b))
            bld = gsdf.Builder(panic_on_error=False)
            s = f.TextLine(bld, "Ab1~c8p23")
            s.Bounds()
            bld.flatten(s)
            parsed += 1
        except Exception:
            failed += 1
    assert parsed > 50 and failed > 50
