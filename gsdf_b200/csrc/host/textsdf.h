// textsdf.h -- host-side mirror of the reference's forge/textsdf package (font.go): TrueType outlines -> polygons ->
// Union2D / Difference2D tree of poly2D nodes (BASELINE config 5, examples/image-text/text.go:24-34).
//
// Scene CONSTRUCTION code (runs once on the host); the hot path evaluates the tree it produces.
//
// The reference delegates font parsing to golang.org/x/image v0.22.0 (go.mod:13; font/sfnt + math/fixed) and curve
// sampling to github.com/soypat/geometry (ms2.Spline3Sampler) -- neither is vendored under the reference tree. Both
// are restated here from their published behaviour (sfnt: cmap format 4/6/12 lookup, hmtx advances, head bounds,
// glyf contour walk with implied on-curve midpoints, Y flipped; Spline3Sampler: bisection to a chord tolerance).
// The polygon vertices they yield are INPUTS to both the CUDA kernels and the CPU oracle ("parity unpinned",
// DESIGN.md); tests pin the outline decoding against FreeType rasterisations of the same font instead.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "builder.h"

namespace gsdfhost {
namespace textsdf {

// sfnt.Segment (x/image/font/sfnt/sfnt.go): op + up to three fixed.Point26_6 arguments (raw 26.6 integers).
enum SegOp { SegMoveTo = 0, SegLineTo = 1, SegQuadTo = 2, SegCubeTo = 3 };
struct Segment {
    int op;
    int32_t x[3], y[3];
};

// The subset of sfnt.Font the reference calls (font.go:54,99,106,125,194-206,223).
class SFNT {
public:
    bool Parse(const uint8_t *data, size_t n, std::string &err);       // sfnt.Parse
    int UnitsPerEm() const { return upm_; }                             // Font.UnitsPerEm
    int NumGlyphs() const { return nglyphs_; }
    // Font.GlyphIndex: 0 (the .notdef glyph) when the rune is not mapped.
    int GlyphIndex(uint32_t rune) const;
    // Font.GlyphAdvance(idx, ppem, HintingNone) in 26.6 units at ppem.
    bool GlyphAdvance(int idx, int32_t ppem, int32_t &adv, std::string &err) const;
    // Font.Kern(i0, i1, ppem, HintingNone): false = ErrNotFound (no `kern` table / no pair). GPOS is not consulted.
    bool Kern(int i0, int i1, int32_t ppem, int32_t &kern) const;
    // Font.Bounds(ppem, HintingNone): {min.x, min.y, max.x, max.y} with Y flipped (min.y = -yMax).
    void Bounds(int32_t ppem, int32_t out[4]) const;
    // Font.LoadGlyph(idx, ppem, nil): contour segments, Y increasing downward.
    bool LoadGlyph(int idx, int32_t ppem, std::vector<Segment> &segs, std::string &err, int depth = 0) const;

private:
    struct Table { uint32_t off = 0, len = 0; };
    bool table(const char *tag, Table &t) const;
    // Every multi-byte read goes through here and is bounds-checked: offsets come from the file, and a damaged font must
    // end in an error or an empty glyph, never in a read outside the buffer (found by fuzzing mutated fonts under ASan).
    uint16_t u16(size_t o) const { return o + 2 <= d_.size() ? (uint16_t)((d_[o] << 8) | d_[o + 1]) : (uint16_t)0; }
    int16_t i16(size_t o) const { return (int16_t)u16(o); }
    uint32_t u32(size_t o) const { return ((uint32_t)u16(o) << 16) | u16(o + 2); }
    bool glyphRange(int idx, uint32_t &beg, uint32_t &end) const;
    int32_t scale(int64_t x, int32_t ppem) const;  // sfnt scale(): round(x*ppem/unitsPerEm), half away from zero

    std::vector<uint8_t> d_;
    std::map<std::string, Table> tabs_;
    int upm_ = 0, nglyphs_ = 0, nhm_ = 0, locaFormat_ = 0;
    int16_t bbox_[4] = {0, 0, 0, 0};
    uint32_t cmapOff_ = 0;
    int cmapFmt_ = 0;
};

// textsdf.Font (font.go:28-38).
class Font {
public:
    // Font.Configure (font.go:40-51): tolerance must be in [0,1); 0 selects the default 0.15.
    bool Configure(float relativeGlyphTolerance, std::string &err);
    bool LoadTTFBytes(const uint8_t *ttf, size_t n, std::string &err);  // font.go:54-62
    // Font.TextLine (font.go:89-141). utf8 text; returns the root node id or -1 with err set.
    NodeId TextLine(Builder &bld, const std::string &utf8, std::string &err);
    NodeId Glyph(Builder &bld, uint32_t rune, std::string &err);        // font.go:159-165
    float Kern(uint32_t c0, uint32_t c1) const;                          // font.go:144-149
    float AdvanceWidth(uint32_t c) const;                                // font.go:152-156
    float scaleout() const;                                              // font.go:208-212
    bool loaded() const { return loaded_; }
    const SFNT &sfnt() const { return sfn_; }

private:
    int32_t scale() const { return (int32_t)sfn_.UnitsPerEm(); }        // font.go:194-197 (ppem = unitsPerEm as raw 26.6)
    NodeId makeGlyph(Builder &bld, uint32_t rune, std::string &err);    // font.go:214-257
    void reset();

    SFNT sfn_;
    bool loaded_ = false;
    float reltol_ = 0.f;
    const Builder *cacheOwner_ = nullptr;
    std::map<uint32_t, NodeId> glyphs_;
};

// font.go:277-330: one contour -> polygon vertices; fill = windingSum < 0.
bool SegmentsToPolygon(const std::vector<Segment> &contour, float tol, float scale, std::vector<Vec2> &poly, bool &fill);
// ms2.Spline3Sampler.SampleBisect for quadratic (c3 unused) / cubic Beziers: appends interior points only.
void SampleBisectQuad(std::vector<Vec2> &dst, Vec2 p0, Vec2 c, Vec2 p1, float tol, int maxDepth);
void SampleBisectCubic(std::vector<Vec2> &dst, Vec2 p0, Vec2 c1, Vec2 c2, Vec2 p1, float tol, int maxDepth);

}  // namespace textsdf
}  // namespace gsdfhost
