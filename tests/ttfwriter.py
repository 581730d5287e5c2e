"""Minimal TrueType (glyf) font writer for tests: turns a dict of glyph outlines into the bytes of a .ttf so that the
C++ sfnt parser (gsdf_b200/csrc/host/textsdf.cpp) can be exercised on any machine, including the GPU box where the
reference's embedded iso-3098.ttf does not exist. Tables written: head hhea maxp hmtx cmap(format 4) loca glyf.

glyphs: {char: {"advance": int, "contours": [[(x, y, on_curve), ...], ...]}} in font units, Y up.
"""
import struct


def _encode_glyph(contours, compact):
    pts = [p for c in contours for p in c]
    if not pts:
        return b""
    xs, ys = [p[0] for p in pts], [p[1] for p in pts]
    out = struct.pack(">hhhhh", len(contours), min(xs), min(ys), max(xs), max(ys))
    end = -1
    for c in contours:
        end += len(c)
        out += struct.pack(">H", end)
    out += struct.pack(">H", 0)  # no instructions
    flags, xb, yb = [], b"", b""
    px = py = 0
    for (x, y, on) in pts:
        f = 1 if on else 0
        dx, dy = x - px, y - py
        px, py = x, y
        if compact and dx == 0:
            f |= 0x10
        elif compact and -255 <= dx <= 255:
            f |= 0x02 | (0x10 if dx > 0 else 0)
            xb += struct.pack("B", abs(dx))
        else:
            xb += struct.pack(">h", dx)
        if compact and dy == 0:
            f |= 0x20
        elif compact and -255 <= dy <= 255:
            f |= 0x04 | (0x20 if dy > 0 else 0)
            yb += struct.pack("B", abs(dy))
        else:
            yb += struct.pack(">h", dy)
        flags.append(f)
    fb = b""
    i = 0
    while i < len(flags):
        j = i
        while compact and j + 1 < len(flags) and flags[j + 1] == flags[i] and j - i < 255:
            j += 1
        if j > i:
            fb += struct.pack("BB", flags[i] | 0x08, j - i)
        else:
            fb += struct.pack("B", flags[i])
        i = j + 1
    out += fb + xb + yb
    if len(out) % 2:
        out += b"\0"
    return out


def _cmap4(mapping):
    """mapping: sorted list of (codepoint, glyph index); one segment per codepoint run with a constant delta."""
    segs = []
    for cp, gi in mapping:
        if segs and segs[-1][1] + 1 == cp and segs[-1][2] == (gi - cp) & 0xFFFF:
            segs[-1][1] = cp
        else:
            segs.append([cp, cp, (gi - cp) & 0xFFFF])
    segs.append([0xFFFF, 0xFFFF, 1])
    n = len(segs)
    sr = 1
    es = 0
    while sr * 2 <= n:
        sr *= 2
        es += 1
    body = struct.pack(">HHHH", n * 2, sr * 2, es, n * 2 - sr * 2)
    body += b"".join(struct.pack(">H", s[1]) for s in segs) + struct.pack(">H", 0)
    body += b"".join(struct.pack(">H", s[0]) for s in segs)
    body += b"".join(struct.pack(">H", s[2]) for s in segs)
    body += b"".join(struct.pack(">H", 0) for _ in segs)
    sub = struct.pack(">HHH", 4, 6 + len(body), 0) + body
    return struct.pack(">HH", 0, 1) + struct.pack(">HHI", 3, 1, 12) + sub


def write_ttf(glyphs, units_per_em, bbox, compact=True, long_loca=False):
    chars = sorted(glyphs)
    glyf = b""
    offs = [0, 0]  # glyph 0 = empty .notdef
    adv = [units_per_em // 2]
    for ch in chars:
        glyf += _encode_glyph(glyphs[ch]["contours"], compact)
        offs.append(len(glyf))
        adv.append(int(glyphs[ch]["advance"]))
    ng = len(chars) + 1
    head = struct.pack(">IIIIHHQQhhhhHHhhh", 0x00010000, 0x00010000, 0, 0x5F0F3CF5, 0, units_per_em, 0, 0, bbox[0], bbox[1], bbox[2],
                       bbox[3], 0, 8, 2, 1 if long_loca else 0, 0)
    hhea = struct.pack(">IhhhHhhhhhhhhhhhH", 0x00010000, bbox[3], bbox[1], 0, max(adv), 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, ng)
    maxp = struct.pack(">IH", 0x00010000, ng) + b"\0" * 26
    hmtx = b"".join(struct.pack(">Hh", a, 0) for a in adv)
    loca = b"".join(struct.pack(">I", o) for o in offs) if long_loca else b"".join(struct.pack(">H", o // 2) for o in offs)
    cmap = _cmap4([(ord(ch), i + 1) for i, ch in enumerate(chars)])
    tables = sorted({"head": head, "hhea": hhea, "maxp": maxp, "hmtx": hmtx, "loca": loca, "cmap": cmap, "glyf": glyf or b"\0\0"}.items())
    nt = len(tables)
    out = struct.pack(">IHHHH", 0x00010000, nt, 0, 0, 0)
    off = 12 + 16 * nt
    body = b""
    for tag, data in tables:
        out += struct.pack(">4sIII", tag.encode(), 0, off + len(body), len(data))
        body += data + b"\0" * (-len(data) % 4)
    return out + body
