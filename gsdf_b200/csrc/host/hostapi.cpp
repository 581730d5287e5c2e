// hostapi.cpp -- extern "C" wrapper over the C++ host layer; see include/gsdf_host.h.
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/gsdf_host.h"
#include "builder.h"
#include "flatten.h"
#include "textsdf.h"
#include "threads.h"

using namespace gsdfhost;

struct gsdfh_builder {
    Builder b;
    std::string err;      // last fatal error of a C call
    std::string errjoin;  // storage for gsdfh_builder_err
};
struct gsdfh_font {
    textsdf::Font f;
    std::string err;
};
struct gsdfh_flat {
    Program prog;
    std::vector<uint8_t> blob;
};

namespace {
int32_t failb(gsdfh_builder *b, const std::string &msg) {
    b->err = msg;
    return -1;
}
bool make_threader(gsdfh_builder *b, int kind, float p0, float p1, int ext, threads::Threader &t) {
    if (kind == 0) { t = threads::Threader::ISO(p0, p1, ext != 0); return true; }
    if (kind == 1) {
        if (!threads::Threader::NPTFromNominal(p0, t)) { b->err = "nominal measurement not found"; return false; }  // npt.go:73
        return true;
    }
    if (kind == 2) { t = threads::Threader::UTS(p0, p1, ext != 0); return true; }                           // UTS{D, TPI, Ext}
    if (kind == 3) { t = threads::Threader::Basic(threads::Kind::Acme, p0, p1); return true; }               // Acme{D, P}
    if (kind == 4) { t = threads::Threader::Basic(threads::Kind::ANSIButtress, p0, p1); return true; }       // ANSIButtress{D, P}
    if (kind == 5) { t = threads::Threader::Basic(threads::Kind::PlasticButtress, p0, p1); return true; }    // PlasticButtress{D, P}
    b->err = "unknown thread kind";
    return false;
}
}  // namespace

extern "C" {

gsdfh_builder *gsdfh_builder_new(void) { return new gsdfh_builder(); }
void gsdfh_builder_free(gsdfh_builder *b) { delete b; }
const char *gsdfh_builder_err(gsdfh_builder *b) {
    b->errjoin = b->b.Err();
    if (!b->err.empty()) { if (!b->errjoin.empty()) b->errjoin += "\n"; b->errjoin += b->err; }
    return b->errjoin.c_str();
}
void gsdfh_builder_clear_errors(gsdfh_builder *b) { b->b.ClearErrors(); b->err.clear(); }

int32_t gsdfh_node(gsdfh_builder *hb, int32_t kind, const float *f, int nf, const int32_t *ip, int ni, const int32_t *ch, int nch,
                   const float *aux, int naux) {
    Builder &b = hb->b;
    auto F = [&](int i) { return i < nf ? f[i] : 0.f; };
    auto I = [&](int i) { return i < ni ? ip[i] : 0; };
    auto C = [&](int i) { return i < nch ? ch[i] : -1; };
    auto pts = [&]() { std::vector<Vec2> v; for (int i = 0; i + 1 < naux; i += 2) v.push_back({aux[i], aux[i + 1]}); return v; };
    auto kids = [&]() { return std::vector<NodeId>(ch, ch + nch); };
    switch (kind) {
    case GSDF_N_SPHERE: return b.NewSphere(F(0));
    case GSDF_N_BOX: return b.NewBox(F(0), F(1), F(2), F(3));
    case GSDF_N_CYLINDER: return b.NewCylinder(F(0), F(1), F(2));
    case GSDF_N_HEX: return b.NewHexagonalPrism(F(0), F(1));
    case GSDF_N_TORUS: return b.NewTorus(F(0), F(1));  // (greaterRadius, lesserRadius)
    case GSDF_N_BOXFRAME: return b.NewBoxFrame(F(0), F(1), F(2), F(3));
    case GSDFH_CALL_TRIPRISM: return b.NewTriangularPrism(F(0), F(1));
    case GSDFH_CALL_BOUNDSBOXFRAME: return b.NewBoundsBoxFrame(Box3{{F(0), F(1), F(2)}, {F(3), F(4), F(5)}});
    case GSDF_N_UNION: return b.Union(kids());
    case GSDF_N_DIFF: return b.Difference(C(0), C(1));
    case GSDF_N_INTERSECT: return b.Intersection(C(0), C(1));
    case GSDF_N_XOR: return b.Xor(C(0), C(1));
    case GSDF_N_SMOOTH_UNION: return b.SmoothUnion(F(0), C(0), C(1));
    case GSDF_N_SMOOTH_DIFF: return b.SmoothDifference(F(0), C(0), C(1));
    case GSDF_N_SMOOTH_INTERSECT: return b.SmoothIntersect(F(0), C(0), C(1));
    case GSDF_N_SCALE: return b.Scale(C(0), F(0));
    case GSDF_N_SYMMETRY: return b.Symmetry(C(0), I(0) & 1, I(0) & 2, I(0) & 4);
    case GSDFH_CALL_ROTATE: return b.Rotate(C(0), F(0), Vec3{F(1), F(2), F(3)});
    case GSDFH_CALL_TRANSFORM16: {
        Mat4 m;
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) m.x[r][c] = F(4 * r + c);
        return b.Transform(C(0), m);
    }
    case GSDF_N_TRANSLATE: return b.Translate(C(0), F(0), F(1), F(2));
    case GSDF_N_OFFSET: return b.Offset(C(0), F(0));
    case GSDF_N_ARRAY: return b.Array(C(0), F(0), F(1), F(2), I(0), I(1), I(2));
    case GSDF_N_ELONGATE: return b.Elongate(C(0), F(0), F(1), F(2));
    case GSDF_N_SHELL: return b.Shell(C(0), F(0));
    case GSDF_N_CIRCARRAY: return b.CircularArray(C(0), I(0), I(1));
    case GSDF_N_TWIST: return b.Twist(C(0), F(0));
    case GSDF_N_BOUNDS3: return b.OverloadShader3DBounds(C(0), Box3{{F(0), F(1), F(2)}, {F(3), F(4), F(5)}});
    case GSDF_N_BOUNDS2: return b.OverloadShader2DBounds(C(0), Box2{{F(0), F(1)}, {F(2), F(3)}});
    case GSDF_N_EXTRUDE: return b.Extrude(C(0), F(0));
    case GSDF_N_REVOLVE: return b.Revolve(C(0), F(0));
    case GSDF_N_SCREW: return b.NewScrew(C(0), F(0), F(1), F(2), F(3));  // pitch, lead, length, taper
    case GSDF_N_LINE2D: return b.NewLine2D(F(0), F(1), F(2), F(3), F(4));
    case GSDF_N_LINES2D: return b.NewLines2D(pts(), F(0));
    case GSDF_N_ARC2D: return b.NewArc(F(0), F(1), F(2));
    case GSDF_N_CIRCLE2D: return b.NewCircle(F(0));
    case GSDF_N_EQTRI2D: return b.NewEquilateralTriangle(F(0));
    case GSDF_N_RECT2D: return b.NewRectangle(F(0), F(1));
    case GSDF_N_HEX2D: return b.NewHexagon(F(0));
    case GSDF_N_OCT2D: return b.NewOctagon(F(0));
    case GSDF_N_POLY2D: return b.NewPolygon(pts());
    case GSDF_N_DIAMOND2D: return b.NewDiamond2D(F(0), F(1));
    case GSDF_N_ROUNDX2D: return b.NewRoundedX(F(0), F(1));
    case GSDF_N_ELLIPSE2D: return b.NewEllipse(F(0), F(1));
    case GSDF_N_BEZIERQ2D: return b.NewQuadraticBezier2D(Vec2{F(0), F(1)}, Vec2{F(2), F(3)}, Vec2{F(4), F(5)}, F(6));
    case GSDF_N_UNION2D: return b.Union2D(kids());
    case GSDF_N_DIFF2D: return b.Difference2D(C(0), C(1));
    case GSDF_N_INTERSECT2D: return b.Intersection2D(C(0), C(1));
    case GSDF_N_XOR2D: return b.Xor2D(C(0), C(1));
    case GSDF_N_ARRAY2D: return b.Array2D(C(0), F(0), F(1), I(0), I(1));
    case GSDF_N_OFFSET2D: return b.Offset2D(C(0), F(0));
    case GSDF_N_TRANSLATE2D: return b.Translate2D(C(0), F(0), F(1));
    case GSDF_N_ROTATE2D: return b.Rotate2D(C(0), F(0));
    case GSDF_N_SYMMETRY2D: return b.Symmetry2D(C(0), I(0) & 1, I(0) & 2);
    case GSDF_N_ANNULUS2D: return b.Annulus(C(0), F(0));
    case GSDF_N_CIRCARRAY2D: return b.CircularArray2D(C(0), I(0), I(1));
    case GSDF_N_SCALE2D: return b.Scale2D(C(0), F(0));
    case GSDF_N_TRANSLATEMULTI2D: return b.TranslateMulti2D(C(0), pts());
    case GSDF_N_ELONGATE2D: return b.Elongate2D(C(0), F(0), F(1));
    }
    return failb(hb, "gsdfh_node: unsupported constructor kind " + std::to_string(kind));
}

int gsdfh_is2d(gsdfh_builder *b, int32_t id) { return b->b.is2D(id) ? 1 : 0; }
int gsdfh_bounds3(gsdfh_builder *b, int32_t id, float out[6]) {
    if (!b->b.is3D(id)) return failb(b, "gsdfh_bounds3: not a 3D node");
    Box3 bb = b->b.Bounds3(id);
    out[0] = bb.min.x; out[1] = bb.min.y; out[2] = bb.min.z; out[3] = bb.max.x; out[4] = bb.max.y; out[5] = bb.max.z;
    return 0;
}
int gsdfh_bounds2(gsdfh_builder *b, int32_t id, float out[4]) {
    if (!b->b.is2D(id)) return failb(b, "gsdfh_bounds2: not a 2D node");
    Box2 bb = b->b.Bounds2(id);
    out[0] = bb.min.x; out[1] = bb.min.y; out[2] = bb.max.x; out[3] = bb.max.y;
    return 0;
}

int32_t gsdfh_thread_profile(gsdfh_builder *b, int kind, float p0, float p1, int ext) {
    threads::Threader t;
    if (!make_threader(b, kind, p0, p1, ext, t)) return -1;
    NodeId id = t.Thread(b->b, b->err);
    return id;
}
int32_t gsdfh_screw(gsdfh_builder *b, float length, int kind, float p0, float p1, int ext) {
    threads::Threader t;
    if (!make_threader(b, kind, p0, p1, ext, t)) return -1;
    return threads::Screw(b->b, length, t, b->err);
}
int32_t gsdfh_nut(gsdfh_builder *b, int kind, float p0, float p1, int ext, int style, float tol) {
    threads::Threader t;
    if (!make_threader(b, kind, p0, p1, ext, t)) return -1;
    return threads::Nut(b->b, t, (threads::NutStyle)style, tol, b->err);
}
int32_t gsdfh_bolt(gsdfh_builder *b, int kind, float p0, float p1, int ext, int style, float tol, float total_len, float shank_len) {
    threads::Threader t;
    if (!make_threader(b, kind, p0, p1, ext, t)) return -1;
    return threads::Bolt(b->b, t, (threads::NutStyle)style, tol, total_len, shank_len, b->err);
}
int32_t gsdfh_hexhead(gsdfh_builder *b, float radius, float height, int round_neg, int round_pos) {
    return threads::HexHead(b->b, radius, height, round_neg != 0, round_pos != 0, b->err);
}
int32_t gsdfh_scene(gsdfh_builder *b, const char *name, float param) {
    std::string n = name ? name : "";
    std::string err;
    NodeId id = -1;
    if (n == "npt-flange") id = scenes::NptFlange(b->b, err);
    else if (n == "bolt") id = scenes::Bolt(b->b, err);
    else if (n == "knurled-cylinder") id = scenes::KnurledCylinder(b->b, param > 0 ? param : 20.f, err);
    else if (n == "fibonacci-showerhead") id = scenes::FibonacciShowerhead(b->b, err);
    else return failb(b, "unknown scene: " + n);
    if (id < 0) b->err = err;
    return id;
}

gsdfh_font *gsdfh_font_new(void) { return new gsdfh_font(); }
void gsdfh_font_free(gsdfh_font *f) { delete f; }
const char *gsdfh_font_err(gsdfh_font *f) { return f->err.c_str(); }
int gsdfh_font_configure(gsdfh_font *f, float tol) { f->err.clear(); return f->f.Configure(tol, f->err) ? 0 : GSDF_EINVAL; }
int gsdfh_font_load_ttf(gsdfh_font *f, const void *ttf, size_t n) {
    f->err.clear();
    if (!ttf) { f->err = "nil font data"; return GSDF_EINVAL; }
    return f->f.LoadTTFBytes(static_cast<const uint8_t *>(ttf), n, f->err) ? 0 : GSDF_EINVAL;
}
int32_t gsdfh_font_textline(gsdfh_font *f, gsdfh_builder *b, const char *utf8) {
    f->err.clear();
    return f->f.TextLine(b->b, utf8 ? utf8 : "", f->err);
}
int32_t gsdfh_font_glyph(gsdfh_font *f, gsdfh_builder *b, uint32_t rune) {
    f->err.clear();
    return f->f.Glyph(b->b, rune, f->err);
}
float gsdfh_font_kern(gsdfh_font *f, uint32_t c0, uint32_t c1) { return f->f.loaded() ? f->f.Kern(c0, c1) : 0.f; }
float gsdfh_font_advance_width(gsdfh_font *f, uint32_t c) { return f->f.loaded() ? f->f.AdvanceWidth(c) : 0.f; }
float gsdfh_font_scaleout(gsdfh_font *f) { return f->f.loaded() ? f->f.scaleout() : 0.f; }
int32_t gsdfh_font_glyph_index(gsdfh_font *f, uint32_t rune) { return f->f.loaded() ? f->f.sfnt().GlyphIndex(rune) : -1; }
int32_t gsdfh_font_glyph_segments(gsdfh_font *f, int32_t gi, int32_t *out7, int32_t maxseg) {
    f->err.clear();
    if (!f->f.loaded()) { f->err = "textsdf: no font loaded"; return -1; }
    std::vector<textsdf::Segment> segs;
    if (!f->f.sfnt().LoadGlyph(gi, f->f.sfnt().UnitsPerEm(), segs, f->err)) return -1;
    if (out7)
        for (int32_t i = 0; i < (int32_t)segs.size() && i < maxseg; i++) {
            out7[7 * i] = segs[i].op;
            for (int k = 0; k < 3; k++) { out7[7 * i + 1 + 2 * k] = segs[i].x[k]; out7[7 * i + 2 + 2 * k] = segs[i].y[k]; }
        }
    return (int32_t)segs.size();
}
int gsdfh_font_info(gsdfh_font *f, int32_t info[6]) {
    if (!f->f.loaded()) return GSDF_EINVAL;
    info[0] = f->f.sfnt().UnitsPerEm(); info[1] = f->f.sfnt().NumGlyphs();
    f->f.sfnt().Bounds(f->f.sfnt().UnitsPerEm(), info + 2);
    return 0;
}

int gsdfh_tree(gsdfh_builder *b, const gsdf_tree_node **nodes, int32_t *nnodes, const int32_t **children, int32_t *nchildren,
               const float **aux, int32_t *naux) {
    *nodes = b->b.nodes().data(); *nnodes = (int32_t)b->b.nodes().size();
    *children = b->b.children().data(); *nchildren = (int32_t)b->b.children().size();
    *aux = b->b.aux().data(); *naux = (int32_t)b->b.aux().size();
    return 0;
}

gsdfh_flat *gsdfh_flatten(gsdfh_builder *b, int32_t root) {
    gsdfh_flat *f = new gsdfh_flat();
    std::string err;
    if (!Flatten(b->b, root, f->prog, err)) { b->err = err; delete f; return nullptr; }
    f->blob = f->prog.blob();
    return f;
}
const void *gsdfh_flat_blob(const gsdfh_flat *f, size_t *nbytes) { *nbytes = f->blob.size(); return f->blob.data(); }
const float *gsdfh_flat_aux(const gsdfh_flat *f, size_t *nfloats) { *nfloats = f->prog.aux.size(); return f->prog.aux.data(); }
void gsdfh_flat_info(const gsdfh_flat *f, int32_t info[5]) {
    info[0] = f->prog.dim; info[1] = f->prog.ninstr; info[2] = (int32_t)(f->prog.chunks.size() / 4);
    info[3] = f->prog.dstack; info[4] = f->prog.pstack;
}
void gsdfh_flat_free(gsdfh_flat *f) { delete f; }

}  // extern "C"
