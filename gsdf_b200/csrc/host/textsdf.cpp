// textsdf.cpp -- see textsdf.h. Mirrors forge/textsdf/font.go; the sfnt / Spline3Sampler pieces restate un-vendored
// third-party modules (golang.org/x/image v0.22.0, github.com/soypat/geometry) from their published behaviour.
#include "textsdf.h"

#include <cstdio>
#include <cstring>

namespace gsdfhost {
namespace textsdf {

// ------------------------------------------------------------------------------------------------ sfnt subset
bool SFNT::table(const char *tag, Table &t) const {
    auto it = tabs_.find(tag);
    if (it == tabs_.end()) return false;
    t = it->second;
    return true;
}

bool SFNT::Parse(const uint8_t *data, size_t n, std::string &err) {
    d_.assign(data, data + n);
    tabs_.clear();
    if (n < 12) { err = "sfnt: invalid font: too short"; return false; }
    const uint32_t ver = u32(0);
    if (ver != 0x00010000u && ver != 0x74727565u /* 'true' */) {
        err = ver == 0x4f54544fu ? "sfnt: PostScript (CFF) outlines are not supported" : "sfnt: invalid font: bad version";
        return false;
    }
    const int nt = u16(4);
    if (12 + 16 * (size_t)nt > n) { err = "sfnt: invalid table directory"; return false; }
    for (int i = 0; i < nt; i++) {
        const size_t o = 12 + 16 * (size_t)i;
        Table t;
        t.off = u32(o + 8);
        t.len = u32(o + 12);
        if ((uint64_t)t.off + t.len > n) { err = "sfnt: table out of bounds"; return false; }
        tabs_[std::string(reinterpret_cast<const char *>(&d_[o]), 4)] = t;
    }
    Table head, maxp, hhea, hmtx, loca, glyf, cmap;
    if (!table("head", head) || head.len < 54) { err = "sfnt: missing head table"; return false; }
    if (!table("maxp", maxp) || maxp.len < 6) { err = "sfnt: missing maxp table"; return false; }
    if (!table("hhea", hhea) || hhea.len < 36) { err = "sfnt: missing hhea table"; return false; }
    if (!table("hmtx", hmtx)) { err = "sfnt: missing hmtx table"; return false; }
    if (!table("loca", loca) || !table("glyf", glyf)) { err = "sfnt: missing loca/glyf table"; return false; }
    if (!table("cmap", cmap) || cmap.len < 4) { err = "sfnt: missing cmap table"; return false; }
    upm_ = u16(head.off + 18);
    if (upm_ <= 0) { err = "sfnt: invalid head table: unitsPerEm"; return false; }
    for (int i = 0; i < 4; i++) bbox_[i] = i16(head.off + 36 + 2 * i);
    locaFormat_ = i16(head.off + 50);
    nglyphs_ = u16(maxp.off + 4);
    nhm_ = u16(hhea.off + 34);
    if (nhm_ < 1 || (uint64_t)4 * nhm_ > hmtx.len) { err = "sfnt: invalid hmtx table"; return false; }
    if ((uint64_t)(nglyphs_ + 1) * (locaFormat_ ? 4 : 2) > loca.len) { err = "sfnt: invalid loca table"; return false; }
    // cmap: prefer a Unicode full-repertoire subtable, then Unicode BMP (platform 0 / platform 3 encodings 1, 10)
    const int nsub = u16(cmap.off + 2);
    int best = -1;
    for (int i = 0; i < nsub; i++) {
        const size_t o = cmap.off + 4 + 8 * (size_t)i;
        if (o + 8 > (size_t)cmap.off + cmap.len) break;
        const int pid = u16(o), eid = u16(o + 2);
        const uint32_t off = u32(o + 4);
        if ((uint64_t)off + 4 > cmap.len) continue;
        const int fmt = u16(cmap.off + off);
        if (fmt != 4 && fmt != 6 && fmt != 12) continue;
        int score = -1;
        if (pid == 0) score = (eid >= 4 ? 4 : 2);
        else if (pid == 3 && eid == 10) score = 5;
        else if (pid == 3 && eid == 1) score = 3;
        if (score > best) { best = score; cmapOff_ = cmap.off + off; cmapFmt_ = fmt; }
    }
    if (best < 0) { err = "sfnt: unsupported cmap encoding"; return false; }
    return true;
}

int SFNT::GlyphIndex(uint32_t r) const {
    const size_t c = cmapOff_;
    if (cmapFmt_ == 4) {
        if (r > 0xffff) return 0;
        const int segX2 = u16(c + 6);
        const size_t endO = c + 14, startO = endO + segX2 + 2, deltaO = startO + segX2, rangeO = deltaO + segX2;
        int lo = 0, hi = segX2 / 2;
        while (lo < hi) {  // first segment whose endCode >= r
            const int mid = (lo + hi) / 2;
            if (u16(endO + 2 * mid) < r) lo = mid + 1; else hi = mid;
        }
        if (lo >= segX2 / 2) return 0;
        const uint32_t start = u16(startO + 2 * lo);
        if (r < start) return 0;
        const uint32_t ro = u16(rangeO + 2 * lo);
        if (ro == 0) return (int)((r + u16(deltaO + 2 * lo)) & 0xffff);
        const size_t go = rangeO + 2 * lo + ro + 2 * (r - start);
        if (go + 2 > d_.size()) return 0;
        const uint32_t g = u16(go);
        return g == 0 ? 0 : (int)((g + u16(deltaO + 2 * lo)) & 0xffff);
    }
    if (cmapFmt_ == 6) {
        const uint32_t first = u16(c + 6), cnt = u16(c + 8);
        if (r < first || r >= first + cnt) return 0;
        return u16(c + 10 + 2 * (r - first));
    }
    if (cmapFmt_ == 12) {
        const uint32_t ng = u32(c + 12);
        uint32_t lo = 0, hi = ng;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) / 2;
            if (u32(c + 16 + 12 * (size_t)mid + 4) < r) lo = mid + 1; else hi = mid;
        }
        if (lo >= ng) return 0;
        const size_t g = c + 16 + 12 * (size_t)lo;
        if (r < u32(g)) return 0;
        return (int)(u32(g + 8) + (r - u32(g)));
    }
    return 0;
}

int32_t SFNT::scale(int64_t x, int32_t ppem) const {
    int64_t v = x * ppem;
    if (v >= 0) v += upm_ / 2; else v -= upm_ / 2;
    return (int32_t)(v / upm_);
}

bool SFNT::GlyphAdvance(int idx, int32_t ppem, int32_t &adv, std::string &err) const {
    if (idx < 0 || idx >= nglyphs_) { err = "sfnt: not found"; return false; }
    Table hmtx;
    table("hmtx", hmtx);
    const int i = idx < nhm_ ? idx : nhm_ - 1;
    adv = scale(u16(hmtx.off + 4 * (size_t)i), ppem);
    return true;
}

bool SFNT::Kern(int i0, int i1, int32_t ppem, int32_t &kern) const {
    Table t;
    if (!table("kern", t) || t.len < 4) return false;
    if (u16(t.off) != 0) return false;  // only the Microsoft/OpenType version-0 header
    const int ntab = u16(t.off + 2);
    size_t o = t.off + 4;
    for (int s = 0; s < ntab; s++) {
        if (o + 14 > (size_t)t.off + t.len) return false;
        const uint32_t len = u16(o + 2), cov = u16(o + 4);
        if ((cov >> 8) == 0 && (cov & 1) && !(cov & 4)) {  // format 0, horizontal, not cross-stream
            const int np = u16(o + 6);
            const uint32_t key = ((uint32_t)i0 << 16) | (uint32_t)i1;
            int lo = 0, hi = np;
            while (lo < hi) {
                const int mid = (lo + hi) / 2;
                const uint32_t k = u32(o + 14 + 6 * (size_t)mid);
                if (k < key) lo = mid + 1; else hi = mid;
            }
            if (lo < np && u32(o + 14 + 6 * (size_t)lo) == key) {
                kern = scale(i16(o + 14 + 6 * (size_t)lo + 4), ppem);
                return true;
            }
        }
        o += len;
    }
    return false;
}

void SFNT::Bounds(int32_t ppem, int32_t out[4]) const {
    out[0] = +scale(bbox_[0], ppem);
    out[1] = -scale(bbox_[3], ppem);
    out[2] = +scale(bbox_[2], ppem);
    out[3] = -scale(bbox_[1], ppem);
}

bool SFNT::glyphRange(int idx, uint32_t &beg, uint32_t &end) const {
    Table loca, glyf;
    table("loca", loca);
    table("glyf", glyf);
    if (locaFormat_) { beg = u32(loca.off + 4 * (size_t)idx); end = u32(loca.off + 4 * (size_t)idx + 4); }
    else { beg = 2u * u16(loca.off + 2 * (size_t)idx); end = 2u * u16(loca.off + 2 * (size_t)idx + 2); }
    if (beg > end || end > glyf.len) return false;
    beg += glyf.off;
    end += glyf.off;
    return true;
}

namespace {
struct Pt { int32_t x, y; };
inline Pt midPoint(Pt p, Pt q) { return {(p.x + q.x) / 2, (p.y + q.y) / 2}; }  // Go integer division: toward zero
inline Segment seg1(int op, Pt a) { Segment s{}; s.op = op; s.x[0] = a.x; s.y[0] = a.y; return s; }
inline Segment seg2(Pt a, Pt b) { Segment s{}; s.op = SegQuadTo; s.x[0] = a.x; s.y[0] = a.y; s.x[1] = b.x; s.y[1] = b.y; return s; }
}  // namespace

bool SFNT::LoadGlyph(int idx, int32_t ppem, std::vector<Segment> &segs, std::string &err, int depth) const {
    if (idx < 0 || idx >= nglyphs_) { err = "sfnt: not found"; return false; }
    if (depth > 8) { err = "sfnt: compound glyph recursion too deep"; return false; }
    uint32_t g0, g1;
    if (!glyphRange(idx, g0, g1)) { err = "sfnt: invalid glyph data"; return false; }
    const size_t first = segs.size();
    if (g1 == g0) return true;  // empty glyph (space)
    if (g1 - g0 < 10) { err = "sfnt: invalid glyph data"; return false; }
    const int nc = i16(g0);
    size_t o = g0 + 10;
    if (nc < 0) {
        // compound glyph: components placed by (dx,dy) and an optional 2.14 linear transform
        for (;;) {
            if (o + 4 > g1) { err = "sfnt: invalid glyph data"; return false; }
            const uint32_t fl = u16(o);
            const int comp = u16(o + 2);
            o += 4;
            int32_t dx, dy;
            if (fl & 0x0001) { if (o + 4 > g1) { err = "sfnt: invalid glyph data"; return false; } dx = i16(o); dy = i16(o + 2); o += 4; }
            else { if (o + 2 > g1) { err = "sfnt: invalid glyph data"; return false; } dx = (int8_t)d_[o]; dy = (int8_t)d_[o + 1]; o += 2; }
            if (!(fl & 0x0002)) { err = "sfnt: unsupported compound glyph (point-matched component)"; return false; }
            int32_t m[4] = {1 << 14, 0, 0, 1 << 14};  // a b c d (2.14)
            bool has = false;
            if (fl & 0x0008) { if (o + 2 > g1) { err = "sfnt: invalid glyph data"; return false; } m[0] = m[3] = i16(o); o += 2; has = true; }
            else if (fl & 0x0040) { if (o + 4 > g1) { err = "sfnt: invalid glyph data"; return false; } m[0] = i16(o); m[3] = i16(o + 2); o += 4; has = true; }
            else if (fl & 0x0080) { if (o + 8 > g1) { err = "sfnt: invalid glyph data"; return false; } for (int k = 0; k < 4; k++) m[k] = i16(o + 2 * k); o += 8; has = true; }
            const size_t b0 = segs.size();
            if (!LoadGlyph(comp, upm_, segs, err, depth + 1)) return false;  // unscaled (ppem = unitsPerEm)
            for (size_t j = b0; j < segs.size(); j++)
                for (int k = 0; k < 3; k++) {
                    int64_t X = segs[j].x[k], Y = segs[j].y[k];  // Y is already flipped
                    if (has) {
                        const int64_t nx = (m[0] * X - m[2] * Y) / (1 << 14);
                        const int64_t ny = (-(int64_t)m[1] * X + m[3] * Y) / (1 << 14);
                        X = nx; Y = ny;
                    }
                    segs[j].x[k] = (int32_t)(X + dx);
                    segs[j].y[k] = (int32_t)(Y - dy);
                }
            if (!(fl & 0x0020)) break;
        }
    } else {
        if (o + 2 * (size_t)nc + 2 > g1) { err = "sfnt: invalid glyph data"; return false; }
        std::vector<int> ends(nc);
        for (int i = 0; i < nc; i++) ends[i] = u16(o + 2 * (size_t)i);
        o += 2 * (size_t)nc;
        const int npts = nc ? ends[nc - 1] + 1 : 0;
        const int ninstr = u16(o);
        o += 2 + (size_t)ninstr;
        std::vector<uint8_t> flags(npts);
        for (int i = 0; i < npts;) {
            if (o >= g1) { err = "sfnt: invalid glyph data"; return false; }
            const uint8_t f = d_[o++];
            flags[i++] = f;
            if (f & 0x08) {
                if (o >= g1) { err = "sfnt: invalid glyph data"; return false; }
                int rep = d_[o++];
                while (rep-- > 0 && i < npts) flags[i++] = f;
            }
        }
        std::vector<Pt> pts(npts);
        int32_t v = 0;
        for (int i = 0; i < npts; i++) {
            const uint8_t f = flags[i];
            if (f & 0x02) { if (o >= g1) { err = "sfnt: invalid glyph data"; return false; } const int d = d_[o++]; v += (f & 0x10) ? d : -d; }
            else if (!(f & 0x10)) { if (o + 2 > g1) { err = "sfnt: invalid glyph data"; return false; } v += i16(o); o += 2; }
            pts[i].x = v;
        }
        v = 0;
        for (int i = 0; i < npts; i++) {
            const uint8_t f = flags[i];
            if (f & 0x04) { if (o >= g1) { err = "sfnt: invalid glyph data"; return false; } const int d = d_[o++]; v += (f & 0x20) ? d : -d; }
            else if (!(f & 0x20)) { if (o + 2 > g1) { err = "sfnt: invalid glyph data"; return false; } v += i16(o); o += 2; }
            pts[i].y = -v;  // sfnt: Y increases downward
        }
        // contour walk (x/image/font/sfnt truetype.go glyfIter): implied on-curve midpoints between consecutive
        // off-curve points; a contour that starts off-curve begins at the first on-curve point (or the midpoint of
        // its first two off-curve points) and the skipped control point is consumed when the contour closes.
        int start = 0;
        for (int c = 0; c < nc; c++) {
            const int end = ends[c];
            if (end < start - 1 || end >= npts) { err = "sfnt: invalid glyph data"; return false; }
            bool firstOnValid = false, firstOffValid = false, lastOffValid = false;
            Pt firstOn{0, 0}, firstOff{0, 0}, lastOff{0, 0};
            for (int i = start; i <= end; i++) {
                const Pt p = pts[i];
                const bool on = flags[i] & 0x01;
                if (!firstOnValid) {
                    if (on) { firstOn = p; firstOnValid = true; segs.push_back(seg1(SegMoveTo, p)); }
                    else if (!firstOffValid) { firstOff = p; firstOffValid = true; }
                    else {
                        firstOn = midPoint(firstOff, p); firstOnValid = true;
                        lastOff = p; lastOffValid = true;
                        segs.push_back(seg1(SegMoveTo, firstOn));
                    }
                } else if (!lastOffValid) {
                    if (!on) { lastOff = p; lastOffValid = true; }
                    else segs.push_back(seg1(SegLineTo, p));
                } else {
                    if (!on) { segs.push_back(seg2(lastOff, midPoint(lastOff, p))); lastOff = p; }
                    else { segs.push_back(seg2(lastOff, p)); lastOffValid = false; }
                }
            }
            if (firstOnValid) {  // close the contour
                if (firstOffValid && lastOffValid) { segs.push_back(seg2(lastOff, midPoint(lastOff, firstOff))); lastOffValid = false; }
                if (!firstOffValid && !lastOffValid) segs.push_back(seg1(SegLineTo, firstOn));
                else if (!firstOffValid && lastOffValid) segs.push_back(seg2(lastOff, firstOn));
                else segs.push_back(seg2(firstOff, firstOn));
            }
            start = end + 1;
        }
    }
    if (depth == 0)
        for (size_t j = first; j < segs.size(); j++)
            for (int k = 0; k < 3; k++) { segs[j].x[k] = scale(segs[j].x[k], ppem); segs[j].y[k] = scale(segs[j].y[k], ppem); }
    return true;
}

// ------------------------------------------------------------------------------------------------ curve sampling
namespace {
// ms2.Spline3.Evaluate with the quadratic / cubic Bezier basis: polynomial basis weights b_k(t), then sum b_k * v_k.
inline Vec2 quadAt(float t, Vec2 p0, Vec2 c, Vec2 p1) {
    const float t2 = t * t;
    const float b0 = 1 - 2 * t + t2, b1 = 2 * t - 2 * t2, b2 = t2;
    return {b0 * p0.x + b1 * c.x + b2 * p1.x, b0 * p0.y + b1 * c.y + b2 * p1.y};
}
inline Vec2 cubicAt(float t, Vec2 p0, Vec2 c1, Vec2 c2, Vec2 p1) {
    const float t2 = t * t, t3 = t2 * t;
    const float b0 = 1 - 3 * t + 3 * t2 - t3, b1 = 3 * t - 6 * t2 + 3 * t3, b2 = 3 * t2 - 3 * t3, b3 = t3;
    return {b0 * p0.x + b1 * c1.x + b2 * c2.x + b3 * p1.x, b0 * p0.y + b1 * c1.y + b2 * c2.y + b3 * p1.y};
}
template <class F>
void bisect(std::vector<Vec2> &dst, const F &at, float t0, Vec2 a, float t1, Vec2 b, float tol, int depth) {
    if (depth <= 0) return;
    const float tm = 0.5f * (t0 + t1);
    const Vec2 m = at(tm);
    const Vec2 chordMid = scale(0.5f, add(a, b));
    if (norm(sub(m, chordMid)) <= tol) return;  // flat enough: the chord stands for the curve
    bisect(dst, at, t0, a, tm, m, tol, depth - 1);
    dst.push_back(m);
    bisect(dst, at, tm, m, t1, b, tol, depth - 1);
}
}  // namespace

void SampleBisectQuad(std::vector<Vec2> &dst, Vec2 p0, Vec2 c, Vec2 p1, float tol, int maxDepth) {
    auto at = [&](float t) { return quadAt(t, p0, c, p1); };
    bisect(dst, at, 0.f, at(0.f), 1.f, at(1.f), tol, maxDepth);
}
void SampleBisectCubic(std::vector<Vec2> &dst, Vec2 p0, Vec2 c1, Vec2 c2, Vec2 p1, float tol, int maxDepth) {
    auto at = [&](float t) { return cubicAt(t, p0, c1, c2, p1); };
    bisect(dst, at, 0.f, at(0.f), 1.f, at(1.f), tol, maxDepth);
}

// font.go:332-337
static inline Vec2 fixedToVec(int32_t x, int32_t y, float scale) { return {(float)x * scale, -(float)y * scale}; }

bool SegmentsToPolygon(const std::vector<Segment> &contour, float tol, float scale, std::vector<Vec2> &poly, bool &fill) {
    float windingSum = 0;
    Vec2 prev{0, 0};
    poly.clear();
    for (const Segment &s : contour) {
        switch (s.op) {
        case SegMoveTo: prev = fixedToVec(s.x[0], s.y[0], scale); break;
        case SegLineTo: {
            const Vec2 p = fixedToVec(s.x[0], s.y[0], scale);
            poly.push_back(prev);
            windingSum += (prev.x - p.x) * (prev.y + p.y);
            prev = p;
            break;
        }
        case SegQuadTo: {
            const Vec2 ctrl = fixedToVec(s.x[0], s.y[0], scale), end = fixedToVec(s.x[1], s.y[1], scale);
            poly.push_back(prev);
            SampleBisectQuad(poly, prev, ctrl, end, tol, 4);
            windingSum += (prev.x - end.x) * (prev.y + end.y);
            prev = end;
            break;
        }
        case SegCubeTo: {
            const Vec2 c1 = fixedToVec(s.x[0], s.y[0], scale), c2 = fixedToVec(s.x[1], s.y[1], scale), end = fixedToVec(s.x[2], s.y[2], scale);
            poly.push_back(prev);
            SampleBisectCubic(poly, prev, c1, c2, end, tol, 4);
            windingSum += (prev.x - end.x) * (prev.y + end.y);
            prev = end;
            break;
        }
        }
    }
    fill = windingSum < 0;
    return true;
}

// ------------------------------------------------------------------------------------------------ Font
void Font::reset() {
    glyphs_.clear();
    cacheOwner_ = nullptr;
    if (reltol_ == 0) reltol_ = 0.15f;  // font.go:76-78
}

bool Font::Configure(float tol, std::string &err) {
    if (tol < 0 || tol >= 1 || tol != tol) { err = "invalid RelativeGlyphTolerance"; return false; }
    reset();
    reltol_ = tol;  // font.go:45 (a zero stays zero until the next reset, as in the reference)
    return true;
}

bool Font::LoadTTFBytes(const uint8_t *ttf, size_t n, std::string &err) {
    SFNT f;
    if (!f.Parse(ttf, n, err)) return false;
    reset();
    sfn_ = f;
    loaded_ = true;
    return true;
}

float Font::scaleout() const {
    int32_t bb[4];
    sfn_.Bounds(scale(), bb);
    const float sx = (float)bb[2] - (float)bb[0], sy = (float)bb[3] - (float)bb[1];
    return 1.f / m32::minf(sx, sy);
}

float Font::Kern(uint32_t c0, uint32_t c1) const {
    int32_t k = 0;
    sfn_.Kern(sfn_.GlyphIndex(c0), sfn_.GlyphIndex(c1), scale(), k);
    return (float)k * scaleout();
}

float Font::AdvanceWidth(uint32_t c) const {
    int32_t a = 0;
    std::string e;
    sfn_.GlyphAdvance(sfn_.GlyphIndex(c), scale(), a, e);
    return (float)a * scaleout();
}

NodeId Font::Glyph(Builder &bld, uint32_t rune, std::string &err) {
    if (!loaded_) { err = "textsdf: no font loaded"; return -1; }
    if (cacheOwner_ != &bld) { glyphs_.clear(); cacheOwner_ = &bld; }  // node ids belong to one Builder
    auto it = glyphs_.find(rune);
    if (it != glyphs_.end()) return it->second;
    const NodeId g = makeGlyph(bld, rune, err);
    if (g >= 0) glyphs_[rune] = g;
    return g;
}

NodeId Font::makeGlyph(Builder &bld, uint32_t rune, std::string &err) {
    const int idx = sfn_.GlyphIndex(rune);
    std::vector<Segment> segs;
    if (!sfn_.LoadGlyph(idx, scale(), segs, err)) return -1;
    const float so = scaleout();
    // splitContours (font.go:260-273): every MoveTo starts a contour
    std::vector<std::vector<Segment>> contours;
    for (const Segment &s : segs) {
        if (s.op == SegMoveTo || contours.empty()) contours.emplace_back();
        contours.back().push_back(s);
    }
    if (contours.empty()) { err = "glyph has no contours"; return -1; }
    NodeId shape = -1;
    for (size_t c = 0; c < contours.size(); c++) {
        std::vector<Vec2> poly;
        bool fill = false;
        SegmentsToPolygon(contours[c], reltol_, so, poly, fill);
        if (poly.empty()) { err = "polygon needs at least 3 distinct vertices"; return -1; }
        const NodeId sdf = bld.NewPolygon(poly);
        const std::string berr = bld.Err();
        if (sdf < 0 || !berr.empty()) { err = berr.empty() ? "NewPolygon failed" : berr; return -1; }  // font.go:329 bld.Err()
        if (c == 0) shape = sdf;
        else shape = fill ? bld.Union2D({shape, sdf}) : bld.Difference2D(shape, sdf);
    }
    return shape;
}

namespace {
// Go's `for _, c := range s`: UTF-8 decode, invalid bytes yield U+FFFD and advance by one.
uint32_t nextRune(const std::string &s, size_t &i) {
    const uint8_t b0 = (uint8_t)s[i];
    auto cont = [&](size_t k) { return i + k < s.size() && (((uint8_t)s[i + k]) & 0xc0) == 0x80; };
    if (b0 < 0x80) { i += 1; return b0; }
    if (b0 >= 0xc2 && b0 <= 0xdf && cont(1)) { const uint32_t r = ((b0 & 0x1f) << 6) | ((uint8_t)s[i + 1] & 0x3f); i += 2; return r; }
    if (b0 >= 0xe0 && b0 <= 0xef && cont(1) && cont(2)) {
        const uint32_t r = ((b0 & 0x0f) << 12) | (((uint8_t)s[i + 1] & 0x3f) << 6) | ((uint8_t)s[i + 2] & 0x3f);
        if (r >= 0x800 && !(r >= 0xd800 && r <= 0xdfff)) { i += 3; return r; }
    }
    if (b0 >= 0xf0 && b0 <= 0xf4 && cont(1) && cont(2) && cont(3)) {
        const uint32_t r = ((b0 & 0x07) << 18) | (((uint8_t)s[i + 1] & 0x3f) << 12) | (((uint8_t)s[i + 2] & 0x3f) << 6) | ((uint8_t)s[i + 3] & 0x3f);
        if (r >= 0x10000 && r <= 0x10ffff) { i += 4; return r; }
    }
    i += 1;
    return 0xfffd;
}
// unicode.IsSpace
bool isSpace(uint32_t c) {
    if (c <= 0xff) return c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r' || c == ' ' || c == 0x85 || c == 0xa0;
    return c == 0x1680 || (c >= 0x2000 && c <= 0x200a) || c == 0x2028 || c == 0x2029 || c == 0x202f || c == 0x205f || c == 0x3000;
}
// unicode.IsGraphic restricted to what can be decided without the Unicode tables: control (Cc), format (Cf: soft
// hyphen, zero-width marks, bidi controls, BOM), surrogates and the line / paragraph separators are not graphic.
bool isGraphic(uint32_t c) {
    if (c < 0x20 || (c >= 0x7f && c < 0xa0) || c == 0xad) return false;
    if ((c >= 0x200b && c <= 0x200f) || c == 0x2028 || c == 0x2029 || (c >= 0x202a && c <= 0x202e) || (c >= 0x2060 && c <= 0x206f) || c == 0xfeff) return false;
    if (c >= 0xd800 && c <= 0xdfff) return false;
    return c <= 0x10ffff;
}
std::string quoteRune(uint32_t c) {
    char buf[16];
    if (c >= 0x20 && c < 0x7f) snprintf(buf, sizeof buf, "'%c'", (char)c);
    else snprintf(buf, sizeof buf, "'\\u%04x'", c);
    return buf;
}
}  // namespace

NodeId Font::TextLine(Builder &bld, const std::string &s, std::string &err) {
    if (!loaded_) { err = "textsdf: no font loaded"; return -1; }
    std::vector<NodeId> shapes;
    const int32_t ppem = scale();
    int idxPrev = 0;
    int32_t xOfs = 0;
    const float so = scaleout();
    for (size_t i = 0; i < s.size();) {
        const size_t ic = i;
        const uint32_t c = nextRune(s, i);
        if (!isGraphic(c)) { err = "char " + quoteRune(c) + " not graphic"; return -1; }
        const int idx = sfn_.GlyphIndex(c);
        int32_t advance = 0;
        std::string e;
        if (!sfn_.GlyphAdvance(idx, ppem, advance, e)) { err = "char " + quoteRune(c) + " advance: " + e; return -1; }
        if (isSpace(c)) {
            if (c == '\t') advance *= 4;
            xOfs += advance;
            continue;
        }
        NodeId charshape = Glyph(bld, c, e);
        if (charshape < 0) { err = "char " + quoteRune(c) + ": " + e; return -1; }
        if (ic > 0) {
            int32_t kern = 0;
            if (sfn_.Kern(idxPrev, idx, ppem, kern)) xOfs += kern;
        }
        idxPrev = idx;
        charshape = bld.Translate2D(charshape, (float)xOfs * so, 0);
        shapes.push_back(charshape);
        xOfs += advance;
    }
    if (shapes.size() == 1) return shapes[0];
    if (shapes.empty()) { err = "no text provided"; return -1; }
    return bld.Union2D(shapes);
}

}  // namespace textsdf
}  // namespace gsdfhost
