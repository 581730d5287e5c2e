#!/usr/bin/env python3
"""Regenerates tests/golden/iso3098_subset.json and iso3098_raster.npz from the reference's embedded font
(/root/reference/forge/textsdf/iso-3098.ttf, embed.go:10-16). Run from the repo root in the build container:

    python tests/golden/make_golden_font.py

* iso3098_subset.json: the glyf outlines (control points, on-curve flags, contour ends), advances, unitsPerEm and head
  bbox of the glyphs the reference's own text runs use -- "Abc123~" (examples/image-text/text.go:33) and "Abp8"
  (forge/textsdf/glyph_test.go:14) -- extracted by an independent pure-Python glyf reader (NOT the C++ parser under
  test). tests/ttfwriter.py turns it back into a .ttf so the parser and config 5 run on the GPU box, where
  /root/reference does not exist.
* iso3098_raster.npz: FreeType (PIL) monochrome rasterisations of the same glyphs at 800 px/em, sampled every 8 px. They pin the outline
  decoding (implied on-curve points, contour closing, Y flip, winding -> fill/hole) against an implementation that is
  neither ours nor the reference's: the sign of the polygon SDF must agree with FreeType away from the outline.
"""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FONT = "/root/reference/forge/textsdf/iso-3098.ttf"
CHARS = "Abc123~p8"
PX_PER_EM = 800


def read_font(path):
    d = open(path, "rb").read()
    nt = struct.unpack(">H", d[4:6])[0]
    tabs = {}
    for i in range(nt):
        tag, _, off, ln = struct.unpack(">4sIII", d[12 + 16 * i:28 + 16 * i])
        tabs[tag.decode()] = (off, ln)
    h = tabs["head"][0]
    upm = struct.unpack(">H", d[h + 18:h + 20])[0]
    bbox = list(struct.unpack(">4h", d[h + 36:h + 44]))
    long_loca = struct.unpack(">h", d[h + 50:h + 52])[0]
    ng = struct.unpack(">H", d[tabs["maxp"][0] + 4:tabs["maxp"][0] + 6])[0]
    nhm = struct.unpack(">H", d[tabs["hhea"][0] + 34:tabs["hhea"][0] + 36])[0]
    lo = tabs["loca"][0]
    if long_loca:
        offs = struct.unpack(">%dI" % (ng + 1), d[lo:lo + 4 * (ng + 1)])
    else:
        offs = [2 * v for v in struct.unpack(">%dH" % (ng + 1), d[lo:lo + 2 * (ng + 1)])]
    # cmap format 4, platform 3 encoding 1
    c = tabs["cmap"][0]
    nsub = struct.unpack(">H", d[c + 2:c + 4])[0]
    sub = None
    for i in range(nsub):
        pid, eid, off = struct.unpack(">HHI", d[c + 4 + 8 * i:c + 12 + 8 * i])
        if (pid, eid) == (3, 1):
            sub = c + off
    assert sub is not None and struct.unpack(">H", d[sub:sub + 2])[0] == 4
    segx2 = struct.unpack(">H", d[sub + 6:sub + 8])[0]
    n = segx2 // 2
    ends = struct.unpack(">%dH" % n, d[sub + 14:sub + 14 + segx2])
    starts = struct.unpack(">%dH" % n, d[sub + 16 + segx2:sub + 16 + 2 * segx2])
    deltas = struct.unpack(">%dH" % n, d[sub + 16 + 2 * segx2:sub + 16 + 3 * segx2])
    ro_off = sub + 16 + 3 * segx2
    ros = struct.unpack(">%dH" % n, d[ro_off:ro_off + segx2])

    def gid(cp):
        for k in range(n):
            if starts[k] <= cp <= ends[k]:
                if ros[k] == 0:
                    return (cp + deltas[k]) & 0xFFFF
                a = ro_off + 2 * k + ros[k] + 2 * (cp - starts[k])
                g = struct.unpack(">H", d[a:a + 2])[0]
                return (g + deltas[k]) & 0xFFFF if g else 0
        return 0

    def glyph(gi):
        g0 = tabs["glyf"][0] + offs[gi]
        g1 = tabs["glyf"][0] + offs[gi + 1]
        if g1 == g0:
            return []
        nc = struct.unpack(">h", d[g0:g0 + 2])[0]
        assert nc >= 0, "composite glyph"
        o = g0 + 10
        ends_ = struct.unpack(">%dH" % nc, d[o:o + 2 * nc])
        o += 2 * nc
        ni = struct.unpack(">H", d[o:o + 2])[0]
        o += 2 + ni
        npts = ends_[-1] + 1
        flags = []
        while len(flags) < npts:
            f = d[o]
            o += 1
            flags.append(f)
            if f & 8:
                r = d[o]
                o += 1
                flags.extend([f] * r)
        flags = flags[:npts]
        xs, v = [], 0
        for f in flags:
            if f & 2:
                v += d[o] if f & 0x10 else -d[o]
                o += 1
            elif not f & 0x10:
                v += struct.unpack(">h", d[o:o + 2])[0]
                o += 2
            xs.append(v)
        ys, v = [], 0
        for f in flags:
            if f & 4:
                v += d[o] if f & 0x20 else -d[o]
                o += 1
            elif not f & 0x20:
                v += struct.unpack(">h", d[o:o + 2])[0]
                o += 2
            ys.append(v)
        contours, s = [], 0
        for e in ends_:
            contours.append([[xs[i], ys[i], flags[i] & 1] for i in range(s, e + 1)])
            s = e + 1
        return contours

    def advance(gi):
        a = tabs["hmtx"][0] + 4 * min(gi, nhm - 1)
        return struct.unpack(">H", d[a:a + 2])[0]

    return dict(upm=upm, bbox=bbox, gid=gid, glyph=glyph, advance=advance, has_kern="kern" in tabs or "GPOS" in tabs)


def rasters():
    """Rendered at 800 px/em (hinting moves an outline by at most half a pixel = 0.6 font units there), then sampled at
    every 8th pixel centre: a 128 x 160 grid of inside/outside bits with 10 font units between samples."""
    from PIL import Image, ImageDraw, ImageFont
    font = ImageFont.truetype(FONT, PX_PER_EM)  # size = pixels per em
    out = {}
    W, H, X0, YB, S = 128, 160, 128, 960, 8  # samples; glyph origin at pixel (X0, YB) on the baseline; stride
    for ch in CHARS:
        img = Image.new("1", (W * S, H * S), 0)
        ImageDraw.Draw(img).text((X0, YB), ch, font=font, fill=1, anchor="ls")
        a = np.asarray(img, dtype=np.uint8)[S // 2::S, S // 2::S]
        assert a.shape == (H, W)
        out["u%04x" % ord(ch)] = np.packbits(a, axis=None)
    out["meta"] = np.array([W, H, X0, YB, PX_PER_EM, S], np.int32)
    return out


def text_image():
    """Oracle bits of config 5 at 96 x 24 (frozen regression pin of the text scene + 2-D evaluator path)."""
    root = os.path.dirname(os.path.dirname(HERE))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import fontfix
    from gsdf_b200 import gsdf
    from oracle import oracle as O
    bld = gsdf.Builder()
    s = fontfix.text_scene(bld)
    mn, mx = s.Bounds()
    img = O.Tree.from_shader(s).image_eval2(mn, mx, 96, 24)
    np.savez_compressed(os.path.join(HERE, "text_image.npz"), bounds=np.concatenate([mn, mx]), dist=img)


def main():
    f = read_font(FONT)
    assert not f["has_kern"]
    glyphs = {}
    for ch in CHARS:
        gi = f["gid"](ord(ch))
        assert gi != 0
        glyphs[ch] = dict(advance=f["advance"](gi), contours=f["glyph"](gi))
    fix = dict(source="forge/textsdf/iso-3098.ttf (reference embed.go:10-16), extracted by tests/golden/make_golden_font.py",
               unitsPerEm=f["upm"], bbox=f["bbox"], glyphs=glyphs)
    with open(os.path.join(HERE, "iso3098_subset.json"), "w") as fp:
        json.dump(fix, fp, separators=(",", ":"))
    np.savez_compressed(os.path.join(HERE, "iso3098_raster.npz"), **rasters())
    text_image()
    npts = sum(len(c) for g in glyphs.values() for c in g["contours"])
    print("wrote %d glyphs, %d control points" % (len(glyphs), npts))


if __name__ == "__main__":
    sys.exit(main())
