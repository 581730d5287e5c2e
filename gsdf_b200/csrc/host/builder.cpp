// builder.cpp -- host-side mirror of gsdf.Builder; see builder.h. Citations are to the reference repository.
#include "builder.h"

#include <cstring>

namespace gsdfhost {

namespace {
// gsdf.go:16-25
constexpr double kTribisect = 0.8660254037844386467637231707529361834714026269051903140279034897;
constexpr double kSqrt3 = 1.7320508075688772935274463415058723669428052538103806280558069794;
constexpr float kLargenum = 1e20f;
constexpr float kEpstol = 6e-7f;
inline bool isInfPos(float x) { return std::isinf(x) && x > 0; }
}  // namespace

NodeId Builder::push(int kind, std::initializer_list<float> f, const std::vector<NodeId> &ch, std::initializer_list<int> ip) {
    gsdf_tree_node n;
    std::memset(&n, 0, sizeof n);
    n.kind = kind;
    n.nchild = (int32_t)ch.size();
    n.child_off = (int32_t)children_.size();
    for (NodeId c : ch) children_.push_back(c);
    int i = 0;
    for (float v : f) n.fparam[i++] = v;
    i = 0;
    for (int v : ip) n.iparam[i++] = v;
    n.aux_off = (int32_t)aux_.size();
    n.aux_cnt = 0;
    nodes_.push_back(n);
    return (NodeId)nodes_.size() - 1;
}

bool Builder::is3D(NodeId id) const { return valid(id) && nodes_[id].kind < GSDF_N_LINE2D; }
bool Builder::is2D(NodeId id) const { return valid(id) && nodes_[id].kind >= GSDF_N_LINE2D; }
bool Builder::need3(NodeId s, const char *who) {
    if (is3D(s)) return true;
    shapeErrorf(std::string("nil or non-3D SDF argument: ") + who);  // gsdf.go:108-110 (nilsdf panics in Go)
    return false;
}
bool Builder::need2(NodeId s, const char *who) {
    if (is2D(s)) return true;
    shapeErrorf(std::string("nil or non-2D SDF argument: ") + who);
    return false;
}
std::string Builder::Err() const {
    std::string r;
    for (size_t i = 0; i < errs_.size(); i++) { if (i) r += "\n"; r += errs_[i]; }
    return r;
}

// ------------------------------------------------------------------ 3D primitives (primitives.go)
NodeId Builder::NewSphere(float r) {
    if (!(r > 0)) shapeErrorf("zero or negative sphere radius");  // :29-32
    return push(GSDF_N_SPHERE, {r}, {});
}
NodeId Builder::NewBox(float x, float y, float z, float round) {
    if (round < 0 || round > x / 2 || round > y / 2 || round > z / 2) shapeErrorf("invalid box rounding value");  // :66
    if (x <= 0 || y <= 0 || z <= 0) shapeErrorf("zero or negative box dimension");                               // :69
    return push(GSDF_N_BOX, {x, y, z, round}, {});
}
NodeId Builder::NewCylinder(float r, float h, float rounding) {
    bool okRounding = rounding >= 0 && rounding < r && rounding < h / 2;  // :108
    if (!okRounding) shapeErrorf("invalid cylinder rounding");
    if (!(r > 0 && h > 0)) shapeErrorf("bad cylinder dimension");         // :112
    return push(GSDF_N_CYLINDER, {r, h, rounding}, {});
}
NodeId Builder::NewHexagonalPrism(float face2Face, float h) {
    if (face2Face <= 0 || h <= 0) shapeErrorf("invalid hexagonal prism parameter");  // :158
    return push(GSDF_N_HEX, {face2Face, h}, {});
}
NodeId Builder::NewTriangularPrism(float triHeight, float extrudeLength) {
    if (!(extrudeLength > 0 && !isInfPos(extrudeLength))) shapeErrorf("bad triangular prism extrude length");  // :199
    return Extrude(NewEquilateralTriangle(triHeight), extrudeLength);
}
NodeId Builder::NewTorus(float greaterRadius, float lesserRadius) {
    if (greaterRadius < 2 * lesserRadius) shapeErrorf("too large torus lesser radius");  // :217
    if (greaterRadius <= 0 || lesserRadius <= 0) shapeErrorf("invalid torus parameter");
    return push(GSDF_N_TORUS, {lesserRadius, greaterRadius}, {});
}
NodeId Builder::NewBoxFrame(float dimX, float dimY, float dimZ, float e) {
    e /= 2;  // :255
    if (dimX <= 0 || dimY <= 0 || dimZ <= 0 || e <= 0) shapeErrorf("negative or zero BoxFrame dimension");
    if (2 * e > minComp(Vec3{dimX, dimY, dimZ})) shapeErrorf("BoxFrame edge thickness too large");
    return push(GSDF_N_BOXFRAME, {dimX, dimY, dimZ, e}, {});
}
NodeId Builder::NewBoundsBoxFrame(const Box3 &bb) {  // :12-21
    Vec3 size = bb.size();
    float frameThickness = maxComp(size) / 256;
    size = addScalar(2 * frameThickness, size);
    NodeId bounding = NewBoxFrame(size.x, size.y, size.z, frameThickness);
    Vec3 c = bb.center();
    return Translate(bounding, c.x, c.y, c.z);
}

// ------------------------------------------------------------------ 3D operations (operations.go)
NodeId Builder::Union(const std::vector<NodeId> &shaders) {
    if (shaders.size() < 2) { shapeErrorf("need at least 2 arguments to Union"); return -1; }  // :36 (panics in Go)
    std::vector<NodeId> joined;
    for (NodeId s : shaders) {
        if (!need3(s, "Union")) return -1;
        const gsdf_tree_node &n = nodes_[s];
        if (n.kind == GSDF_N_UNION) {  // :44-48 nested unions are absorbed
            for (int k = 0; k < n.nchild; k++) joined.push_back(children_[n.child_off + k]);
        } else {
            joined.push_back(s);
        }
    }
    return push(GSDF_N_UNION, {}, joined);
}
NodeId Builder::Difference(NodeId a, NodeId b) {
    if (!need3(a, "Difference") || !need3(b, "Difference")) return -1;
    return push(GSDF_N_DIFF, {}, {a, b});
}
NodeId Builder::Intersection(NodeId a, NodeId b) {
    if (!need3(a, "Intersection") || !need3(b, "Intersection")) return -1;
    return push(GSDF_N_INTERSECT, {}, {a, b});
}
NodeId Builder::Xor(NodeId a, NodeId b) {
    if (!need3(a, "Xor") || !need3(b, "Xor")) return -1;
    return push(GSDF_N_XOR, {}, {a, b});
}
NodeId Builder::Scale(NodeId s, float f) {
    if (!need3(s, "Scale")) return -1;
    return push(GSDF_N_SCALE, {f}, {s});
}
NodeId Builder::Symmetry(NodeId s, bool mx, bool my, bool mz) {
    if (!need3(s, "Symmetry")) return -1;
    if (!mx && !my && !mz) shapeErrorf("ineffective symmetry");  // :286
    return push(GSDF_N_SYMMETRY, {}, {s}, {(mx ? 1 : 0) | (my ? 2 : 0) | (mz ? 4 : 0)});
}
NodeId Builder::Transform(NodeId s, const Mat4 &m) {
    if (!need3(s, "Transform")) return -1;
    float det = determinant(m);
    if (m32::absf(det) < kEpstol) shapeErrorf("singular Mat4");  // :341-344
    Mat4 inv = inverse(m);
    NodeId id = push(GSDF_N_TRANSFORM,
                     {inv.x[0][0], inv.x[0][1], inv.x[0][2], inv.x[0][3], inv.x[1][0], inv.x[1][1], inv.x[1][2], inv.x[1][3],
                      inv.x[2][0], inv.x[2][1], inv.x[2][2], inv.x[2][3]},
                     {s});
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) aux_.push_back(m.x[r][c]);
    nodes_[id].aux_cnt = 12;
    return id;
}
NodeId Builder::Rotate(NodeId s, float radians, Vec3 axis) {
    if (axis.x == 0 && axis.y == 0 && axis.z == 0) shapeErrorf("null vector");  // :395
    return Transform(s, rotationMat4(radians, axis));
}
NodeId Builder::Translate(NodeId s, float dx, float dy, float dz) {
    if (!need3(s, "Translate")) return -1;
    return push(GSDF_N_TRANSLATE, {dx, dy, dz}, {s});
}
NodeId Builder::Offset(NodeId s, float sdfAdd) {
    if (!need3(s, "Offset")) return -1;
    return push(GSDF_N_OFFSET, {sdfAdd}, {s});
}
NodeId Builder::Array(NodeId s, float sx, float sy, float sz, int nx, int ny, int nz) {
    if (!need3(s, "Array")) return -1;
    if (nx <= 0 || ny <= 0 || nz <= 0) shapeErrorf("invalid array repeat param");  // :489
    if (sx <= 0 || sy <= 0 || sz <= 0) shapeErrorf("invalid array spacing");
    return push(GSDF_N_ARRAY, {sx, sy, sz}, {s}, {nx, ny, nz});
}
NodeId Builder::SmoothUnion(float k, NodeId a, NodeId b) {
    if (!need3(a, "SmoothUnion") || !need3(b, "SmoothUnion")) return -1;
    return push(GSDF_N_SMOOTH_UNION, {k}, {a, b});
}
NodeId Builder::SmoothDifference(float k, NodeId a, NodeId b) {
    if (!need3(a, "SmoothDifference") || !need3(b, "SmoothDifference")) return -1;
    return push(GSDF_N_SMOOTH_DIFF, {k}, {a, b});
}
NodeId Builder::SmoothIntersect(float k, NodeId a, NodeId b) {
    if (!need3(a, "SmoothIntersect") || !need3(b, "SmoothIntersect")) return -1;
    return push(GSDF_N_SMOOTH_INTERSECT, {k}, {a, b});
}
NodeId Builder::Elongate(NodeId s, float dx, float dy, float dz) {
    if (!need3(s, "Elongate")) return -1;
    return push(GSDF_N_ELONGATE, {dx, dy, dz}, {s});
}
NodeId Builder::Shell(NodeId s, float thickness) {
    if (!need3(s, "Shell")) return -1;
    return push(GSDF_N_SHELL, {thickness}, {s});
}
NodeId Builder::CircularArray(NodeId s, int numInstances, int circleDiv) {
    if (!need3(s, "CircularArray")) return -1;
    if (circleDiv <= 1 || numInstances <= 0) shapeErrorf("invalid circarray repeat param");  // :768
    if (numInstances > circleDiv) shapeErrorf("bad circular array instances, must be less than or equal to circleDiv");
    return push(GSDF_N_CIRCARRAY, {}, {s}, {numInstances, circleDiv});
}
NodeId Builder::Twist(NodeId s, float k) {
    if (!need3(s, "Twist")) return -1;
    if (k == 0) shapeErrorf("zero twist parameter");  // :839
    return push(GSDF_N_TWIST, {k}, {s});
}

NodeId Builder::OverloadShader3DBounds(NodeId s, const Box3 &bb) {
    if (!need3(s, "OverloadShader3DBounds")) return -1;
    return push(GSDF_N_BOUNDS3, {bb.min.x, bb.min.y, bb.min.z, bb.max.x, bb.max.y, bb.max.z}, {s});
}
NodeId Builder::OverloadShader2DBounds(NodeId s, const Box2 &bb) {
    if (!need2(s, "OverloadShader2DBounds")) return -1;
    return push(GSDF_N_BOUNDS2, {bb.min.x, bb.min.y, bb.max.x, bb.max.y}, {s});
}

// ------------------------------------------------------------------ 2D -> 3D
NodeId Builder::Extrude(NodeId s2, float h) {
    if (!need2(s2, "Extrude")) return -1;
    if (h < 0) shapeErrorf("bad extrusion length");  // operations2d.go:108
    return push(GSDF_N_EXTRUDE, {h}, {s2});
}
NodeId Builder::Revolve(NodeId s2, float axisOffset) {
    if (!need2(s2, "Revolve")) return -1;
    if (axisOffset < 0) shapeErrorf("negative axis offset");  // operations2d.go:153
    return push(GSDF_N_REVOLVE, {axisOffset}, {s2});
}
NodeId Builder::NewScrew(NodeId thread2d, float pitch, float lead, float length, float taper) {
    if (!need2(thread2d, "Screw")) return -1;
    if (length <= 0) { shapeErrorf("need greater than zero length"); return -1; }  // threads.go:80
    return push(GSDF_N_SCREW, {pitch, lead, length / 2, taper}, {thread2d});
}

// ------------------------------------------------------------------ 2D primitives (primitives2d.go)
NodeId Builder::NewCircle(float r) {
    if (!(r > 0 && !isInfPos(r))) shapeErrorf("bad circle radius");  // :228
    return push(GSDF_N_CIRCLE2D, {r}, {});
}
NodeId Builder::NewLine2D(float x0, float y0, float x1, float y1, float width) {
    bool hasNaN = std::isnan(x0) || std::isnan(y0) || std::isnan(x1) || std::isnan(y1) || std::isnan(width);  // :15
    if (hasNaN) shapeErrorf("NaN argument to NewLine2D");
    else if (width < 0) shapeErrorf("negative thickness to NewLine2D");
    Vec2 a{x0, y0}, b{x1, y1};
    float lineLen = norm(sub(a, b));
    if (lineLen < width * 1e-6f || lineLen < kEpstol) {  // :23-28 degenerate line -> circle
        if (width == 0) shapeErrorf("infimal line");
        return NewCircle(width / 2);
    }
    return push(GSDF_N_LINE2D, {width, x0, y0, x1, y1}, {});
}
NodeId Builder::NewLines2D(const std::vector<Vec2> &pts, float width) {
    size_t nseg = pts.size() / 2;
    if (width < 0) shapeErrorf("negative thickness to NewLines2D");  // :63
    if (nseg < 2) shapeErrorf("empty or single points");
    for (size_t i = 0; i + 1 < nseg; i++)
        if (pts[2 * i] == pts[2 * i + 1]) shapeErrorf("superimposed points in NewLines2D");
    NodeId id = push(GSDF_N_LINES2D, {width}, {});
    for (size_t i = 0; i < 2 * nseg; i++) { aux_.push_back(pts[i].x); aux_.push_back(pts[i].y); }
    nodes_[id].aux_cnt = (int32_t)(4 * nseg);
    return id;
}
NodeId Builder::NewArc(float radius, float arcAngle, float thick) {
    const float twoPi = m32::kTwoPiF;
    bool ok = radius > 0 && arcAngle > 0 && thick >= 0;  // :170
    if (!ok) shapeErrorf("invalid argument to NewArc2D");
    if (arcAngle > twoPi) shapeErrorf("arc angle exceeds full circle");
    else if (twoPi - arcAngle < kEpstol) arcAngle = twoPi - 1e-7f;
    return push(GSDF_N_ARC2D, {radius, arcAngle, thick}, {});
}
NodeId Builder::NewEquilateralTriangle(float h) {
    if (!(h > 0 && !isInfPos(h))) shapeErrorf("bad equilateral triangle height");  // :266
    return push(GSDF_N_EQTRI2D, {h}, {});
}
NodeId Builder::NewRectangle(float x, float y) {
    if (!(x > 0 && y > 0 && !isInfPos(x) && !isInfPos(y))) shapeErrorf("bad rectangle dimension");  // :308
    return push(GSDF_N_RECT2D, {x, y}, {});
}
NodeId Builder::NewHexagon(float side) {
    if (!(side > 0 && !isInfPos(side))) shapeErrorf("bad hexagon dimension");  // :349
    return push(GSDF_N_HEX2D, {side}, {});
}
NodeId Builder::NewOctagon(float c) {
    if (!(c > 0)) shapeErrorf("bad octagon dimension");  // :386
    return push(GSDF_N_OCT2D, {c}, {});
}
NodeId Builder::NewPolygon(std::vector<Vec2> v) {
    // validatePolygon :471-492
    if (v.empty()) { shapeErrorf("polygon needs at least 3 distinct vertices"); return -1; }
    size_t prev = v.size() - 1;
    if (v[0] == v[prev]) { v.pop_back(); if (!v.empty()) prev = v.size() - 1; }
    if (v.size() < 3) { shapeErrorf("polygon needs at least 3 distinct vertices"); return -1; }
    for (size_t i = 0; i < v.size(); i++) {
        if (std::isnan(v[i].x) || std::isnan(v[i].y)) { shapeErrorf("NaN value in vertices"); break; }
        if (v[i] == v[prev]) { shapeErrorf("found two consecutive equal vertices in polygon"); break; }
        prev = i;
    }
    NodeId id = push(GSDF_N_POLY2D, {}, {});
    for (Vec2 p : v) { aux_.push_back(p.x); aux_.push_back(p.y); }
    nodes_[id].aux_cnt = (int32_t)(2 * v.size());
    return id;
}
NodeId Builder::NewDiamond2D(float w, float h) {
    if (!(w > 0 && h > 0 && !isInfPos(w) && !isInfPos(h))) shapeErrorf("bad diamond dimension");  // :561
    return push(GSDF_N_DIAMOND2D, {w, h}, {});
}
NodeId Builder::NewRoundedX(float width, float thick) {
    if (!(width > 0 && thick > 0 && !isInfPos(width) && !isInfPos(thick))) shapeErrorf("bad x dimension");  // :603
    return push(GSDF_N_ROUNDX2D, {width, thick}, {});
}

NodeId Builder::NewEllipse(float a, float b) {
    if (!(a > 0 && b > 0 && !isInfPos(a) && !isInfPos(b))) shapeErrorf("bad ellipse dimension");  // :422
    return push(GSDF_N_ELLIPSE2D, {a, b}, {});
}
NodeId Builder::NewQuadraticBezier2D(Vec2 a, Vec2 b, Vec2 c, float thick) {  // :644 (no validation in the reference)
    return push(GSDF_N_BEZIERQ2D, {a.x, a.y, b.x, b.y, c.x, c.y, thick}, {});
}

// ------------------------------------------------------------------ 2D operations (operations2d.go)
NodeId Builder::Union2D(const std::vector<NodeId> &shaders) {
    if (shaders.size() < 2) { shapeErrorf("need at least 2 arguments to Union2D"); return -1; }
    std::vector<NodeId> joined;
    for (NodeId s : shaders) {
        if (!need2(s, "Union2D")) return -1;
        const gsdf_tree_node &n = nodes_[s];
        if (n.kind == GSDF_N_UNION2D) {
            for (int k = 0; k < n.nchild; k++) joined.push_back(children_[n.child_off + k]);
        } else {
            joined.push_back(s);
        }
    }
    return push(GSDF_N_UNION2D, {}, joined);
}
NodeId Builder::Difference2D(NodeId a, NodeId b) {
    if (!need2(a, "Difference2D") || !need2(b, "Difference2D")) return -1;
    return push(GSDF_N_DIFF2D, {}, {a, b});
}
NodeId Builder::Intersection2D(NodeId a, NodeId b) {
    if (!need2(a, "Intersection2D") || !need2(b, "Intersection2D")) return -1;
    return push(GSDF_N_INTERSECT2D, {}, {a, b});
}
NodeId Builder::Xor2D(NodeId a, NodeId b) {
    if (!need2(a, "Xor2D") || !need2(b, "Xor2D")) return -1;
    return push(GSDF_N_XOR2D, {}, {a, b});
}
NodeId Builder::Array2D(NodeId s, float sx, float sy, int nx, int ny) {
    if (!need2(s, "Array2D")) return -1;
    if (nx <= 0 || ny <= 0) shapeErrorf("invalid array repeat param");  // :333
    if (!(sx > 0 && sy > 0 && !isInfPos(sx) && !isInfPos(sy))) shapeErrorf("bad array spacing");
    return push(GSDF_N_ARRAY2D, {sx, sy}, {s}, {nx, ny});
}
NodeId Builder::Offset2D(NodeId s, float f) {
    if (!need2(s, "Offset2D")) return -1;
    return push(GSDF_N_OFFSET2D, {f}, {s});
}
NodeId Builder::Translate2D(NodeId s, float dx, float dy) {
    if (!need2(s, "Translate2D")) return -1;
    return push(GSDF_N_TRANSLATE2D, {dx, dy}, {s});
}
NodeId Builder::Rotate2D(NodeId s, float theta) {
    if (!need2(s, "Rotate2D")) return -1;
    Mat2 m = rotationMat2(theta);
    if (m32::absf(determinant(m)) < kEpstol) shapeErrorf("badly conditioned rotation");  // :496
    Mat2 inv = inverse(m);
    return push(GSDF_N_ROTATE2D, {inv.x00, inv.x01, inv.x10, inv.x11, m.x00, m.x01, m.x10, m.x11}, {s});
}
NodeId Builder::Symmetry2D(NodeId s, bool mx, bool my) {
    if (!need2(s, "Symmetry2D")) return -1;
    if (!mx && !my) shapeErrorf("ineffective symmetry");
    return push(GSDF_N_SYMMETRY2D, {}, {s}, {(mx ? 1 : 0) | (my ? 2 : 0)});
}
NodeId Builder::Annulus(NodeId s, float sub) {
    if (!need2(s, "Annulus")) return -1;
    if (sub <= 0) shapeErrorf("invalid annular parameter");  // :608
    return push(GSDF_N_ANNULUS2D, {sub}, {s});
}
NodeId Builder::CircularArray2D(NodeId s, int numInstances, int circleDiv) {
    if (!need2(s, "CircularArray2D")) return -1;
    if (circleDiv <= 1 || numInstances <= 0) shapeErrorf("invalid circarray repeat param");
    if (numInstances > circleDiv) shapeErrorf("bad circular array instances, must be less than or equal to circleDiv");
    return push(GSDF_N_CIRCARRAY2D, {}, {s}, {numInstances, circleDiv});
}
NodeId Builder::Scale2D(NodeId s, float f) {
    if (!need2(s, "Scale2D")) return -1;
    return push(GSDF_N_SCALE2D, {f}, {s});
}
NodeId Builder::TranslateMulti2D(NodeId s, const std::vector<Vec2> &disp) {
    if (!need2(s, "TranslateMulti2D")) return -1;
    if (disp.empty()) { shapeErrorf("TranslateMulti2D needs at least one displacement"); return -1; }
    NodeId id = push(GSDF_N_TRANSLATEMULTI2D, {}, {s});
    for (Vec2 d : disp) { aux_.push_back(d.x); aux_.push_back(d.y); }
    nodes_[id].aux_cnt = (int32_t)(2 * disp.size());
    return id;
}
NodeId Builder::Elongate2D(NodeId s, float dx, float dy) {
    if (!need2(s, "Elongate2D")) return -1;
    return push(GSDF_N_ELONGATE2D, {dx, dy}, {s});
}

// ------------------------------------------------------------------ Bounds()
Box3 Builder::Bounds3(NodeId id) const {
    if (!is3D(id)) return Box3{};
    const gsdf_tree_node &n = nodes_[id];
    const float *f = n.fparam;
    auto ch = [&](int k) { return children_[n.child_off + k]; };
    switch (n.kind) {
    case GSDF_N_SPHERE: return {{-f[0], -f[0], -f[0]}, {f[0], f[0], f[0]}};  // primitives.go:57
    case GSDF_N_BOX:
    case GSDF_N_BOXFRAME: return centeredBox(Vec3{}, Vec3{f[0], f[1], f[2]});  // :101, :288
    case GSDF_N_CYLINDER: return {{-f[0], -f[0], -f[1] / 2}, {f[0], f[0], f[1] / 2}};  // :125
    case GSDF_N_HEX: {  // :169
        float l = f[0], lx = l / (float)kTribisect;
        return {{-lx, -l, -f[1]}, {lx, l, f[1]}};
    }
    case GSDF_N_TORUS: {  // :241 (rLesser, rGreater)
        float R = f[0] + f[1];
        return {{-R, -R, -f[0]}, {R, R, f[0]}};
    }
    case GSDF_N_UNION: {  // operations.go:56
        Box3 bb = Bounds3(ch(0));
        for (int k = 1; k < n.nchild; k++) bb = bb.unionWith(Bounds3(ch(k)));
        return bb;
    }
    case GSDF_N_DIFF:
    case GSDF_N_SMOOTH_DIFF: return Bounds3(ch(0));  // :128 (smoothDiff embeds diff :618)
    case GSDF_N_INTERSECT:
    case GSDF_N_SMOOTH_INTERSECT: return Bounds3(ch(0)).intersect(Bounds3(ch(1)));  // :171
    case GSDF_N_XOR:
    case GSDF_N_SMOOTH_UNION: return Bounds3(ch(0)).unionWith(Bounds3(ch(1)));  // :216, :575
    case GSDF_N_SCALE: return Bounds3(ch(0)).scaleOrigin(Vec3{f[0], f[0], f[0]});  // :257
    case GSDF_N_SYMMETRY: {  // :297
        Box3 b = Bounds3(ch(0));
        if (n.iparam[0] & 1) b.min.x = m32::minf(b.min.x, -b.max.x);
        if (n.iparam[0] & 2) b.min.y = m32::minf(b.min.y, -b.max.y);
        if (n.iparam[0] & 4) b.min.z = m32::minf(b.min.z, -b.max.z);
        return b;
    }
    case GSDF_N_TRANSFORM: {  // :362
        Mat4 t;
        const float *a = &aux_[n.aux_off];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) t.x[r][c] = a[4 * r + c];
        return mulBox(t, Bounds3(ch(0)));
    }
    case GSDF_N_TRANSLATE: return Bounds3(ch(0)).addVec(Vec3{f[0], f[1], f[2]});  // :412
    case GSDF_N_OFFSET: {  // :455
        Box3 bb = Bounds3(ch(0));
        bb.max = addScalar(-f[0], bb.max);
        bb.min = addScalar(f[0], bb.min);
        return bb.canon();
    }
    case GSDF_N_ARRAY: {  // :504
        Box3 bb = Bounds3(ch(0));
        Vec3 size = mulElem(Vec3{(float)n.iparam[0], (float)n.iparam[1], (float)n.iparam[2]}, Vec3{f[0], f[1], f[2]});
        bb.max = add(bb.max, size);
        return bb;
    }
    case GSDF_N_ELONGATE: {  // :688
        Box3 b = Bounds3(ch(0));
        b.max = maxElem(b.max, Vec3{});
        b.max = add(b.max, scale(0.5f, Vec3{f[0], f[1], f[2]}));
        b.min = scale(-1, b.max);
        return b;
    }
    case GSDF_N_SHELL: return Bounds3(ch(0));  // :732
    case GSDF_N_BOUNDS3: return {{f[0], f[1], f[2]}, {f[3], f[4], f[5]}};  // glbuild.go:1092
    case GSDF_N_CIRCARRAY: {  // :783
        Box3 bb = Bounds3(ch(0));
        Box2 bb2{{bb.min.x, bb.min.y}, {bb.max.x, bb.max.y}};
        Vec2 verts[4];
        bb2.vertices(verts);
        float angle = m32::kTwoPiF / (float)n.iparam[1];
        Mat2 m = rotationMat2(angle);
        for (int i = 0; i < n.iparam[0] - 1; i++)
            for (int v = 0; v < 4; v++) { verts[v] = mulMatVec(m, verts[v]); bb2 = bb2.includePoint(verts[v]); }
        bb.max.x = bb2.max.x; bb.max.y = bb2.max.y;
        bb.min.x = bb2.min.x; bb.min.y = bb2.min.y;
        return bb;
    }
    case GSDF_N_TWIST: {  // :850
        Box3 bb = Bounds3(ch(0));
        Vec3 vs[8];
        bb.vertices(vs);
        float maxR = 0;
        for (auto &v : vs) { float r = m32::hypot32(v.x, v.y); if (r > maxR) maxR = r; }
        return {{-maxR, -maxR, bb.min.z}, {maxR, maxR, bb.max.z}};
    }
    case GSDF_N_EXTRUDE: {  // operations2d.go:119
        Box2 b2 = Bounds2(ch(0));
        float hd2 = f[0] / 2;
        return {{b2.min.x, b2.min.y, -hd2}, {b2.max.x, b2.max.y, hd2}};
    }
    case GSDF_N_REVOLVE: {  // operations2d.go:168
        Box2 b2 = Bounds2(ch(0));
        float radius = m32::maxf(0, b2.max.x - f[0]);
        return {{-radius, b2.min.y, -radius}, {radius, b2.max.y, radius}};
    }
    case GSDF_N_SCREW: {  // threads.go:184-196
        float r = Bounds2(ch(0)).max.y;
        r += f[2] * m32::tan(f[3]);
        return {{-r, -r, -f[2]}, {r, r, f[2]}};
    }
    }
    return Box3{};
}

Box2 Builder::Bounds2(NodeId id) const {
    if (!is2D(id)) return Box2{};
    const gsdf_tree_node &n = nodes_[id];
    const float *f = n.fparam;
    const float *a = n.aux_cnt ? &aux_[n.aux_off] : nullptr;
    auto ch = [&](int k) { return children_[n.child_off + k]; };
    switch (n.kind) {
    case GSDF_N_LINE2D: {  // primitives2d.go:38
        float w = f[0] / 2;
        Box2 b = Box2{{f[1], f[2]}, {f[3], f[4]}}.canon();
        return {{b.min.x - w, b.min.y - w}, {b.max.x + w, b.max.y + w}};
    }
    case GSDF_N_LINES2D: {  // :98
        float w = f[0] / 2;
        Box2 bb = Box2{{a[0], a[1]}, {a[2], a[3]}};  // ms2.NewBox(x0,y0,x1,y1) canonicalises
        bb = bb.canon();
        for (int i = 4; i + 3 < n.aux_cnt; i += 4) { bb = bb.includePoint({a[i], a[i + 1]}); bb = bb.includePoint({a[i + 2], a[i + 3]}); }
        return {{bb.min.x - w, bb.min.y - w}, {bb.max.x + w, bb.max.y + w}};
    }
    case GSDF_N_ARC2D: {  // :195
        float r = f[0] + f[2];
        float rcos = f[0] * m32::cos(f[1] / 2) - f[2];
        return {{-r, rcos}, {r, r}};
    }
    case GSDF_N_CIRCLE2D: return {{-f[0], -f[0]}, {f[0], f[0]}};  // :235
    case GSDF_N_EQTRI2D: {  // :273
        float side = f[0] / (float)kTribisect;
        float longBisect = side / (float)kSqrt3;
        float shortBisect = longBisect / 2;
        return {{-side / 2, -shortBisect}, {side / 2, longBisect}};
    }
    case GSDF_N_RECT2D:
    case GSDF_N_DIAMOND2D: return {{-(f[0] / 2), -(f[1] / 2)}, {f[0] / 2, f[1] / 2}};  // :315, :568
    case GSDF_N_HEX2D: {  // :356
        float w = f[0] / (float)kTribisect;
        return {{-w, -f[0]}, {w, f[0]}};
    }
    case GSDF_N_OCT2D: return {{-f[0], -f[0]}, {f[0], f[0]}};  // :393
    case GSDF_N_ELLIPSE2D: return {{-f[0], -f[1]}, {f[0], f[1]}};  // :429
    case GSDF_N_POLY2D: {  // :494
        Vec2 mn{kLargenum, kLargenum}, mx{-kLargenum, -kLargenum};
        for (int i = 0; i + 1 < n.aux_cnt; i += 2) { mn = minElem(mn, {a[i], a[i + 1]}); mx = maxElem(mx, {a[i], a[i + 1]}); }
        return {mn, mx};
    }
    case GSDF_N_BEZIERQ2D: {  // :648-672 (iquilezles.org/articles/bezierbbox)
        Vec2 p0{f[0], f[1]}, p1{f[2], f[3]}, p2{f[4], f[5]};
        Vec2 mn = minElem(p0, p2), mx = maxElem(p0, p2);
        if (p1.x < mn.x || p1.x > mx.x || p1.y < mn.y || p1.y > mx.y) {
            Vec2 denom = add(p0, sub(p2, scale(2, p1)));
            Vec2 num = sub(p0, p1);
            Vec2 t{m32::clampf(num.x / denom.x, 0.f, 1.f), m32::clampf(num.y / denom.y, 0.f, 1.f)};
            Vec2 s_{1 - t.x, 1 - t.y};
            Vec2 q1{s_.x * s_.x * p0.x, s_.y * s_.y * p0.y};
            Vec2 q2{2 * (s_.x * t.x * p1.x), 2 * (s_.y * t.y * p1.y)};
            Vec2 q3{p2.x * (t.x * t.x), p2.y * (t.y * t.y)};
            Vec2 q = add(q1, add(q2, q3));
            mn = minElem(mn, q); mx = maxElem(mx, q);
        }
        float h = f[6] / 2;
        return {{mn.x + -h, mn.y + -h}, {mx.x + h, mx.y + h}};
    }
    case GSDF_N_ROUNDX2D: {  // :610
        float xd2 = f[0] / 2 + f[1];
        return {{-xd2, -xd2}, {xd2, xd2}};
    }
    case GSDF_N_UNION2D: {  // operations2d.go:36
        Box2 bb = Bounds2(ch(0));
        for (int k = 1; k < n.nchild; k++) bb = bb.unionWith(Bounds2(ch(k)));
        return bb;
    }
    case GSDF_N_DIFF2D: return Bounds2(ch(0));                                 // :213
    case GSDF_N_INTERSECT2D: return Bounds2(ch(0)).intersect(Bounds2(ch(1)));  // :257
    case GSDF_N_XOR2D: return Bounds2(ch(0)).unionWith(Bounds2(ch(1)));        // :301
    case GSDF_N_ARRAY2D: {                                                     // :349
        Box2 bb = Bounds2(ch(0));
        bb.max = add(bb.max, Vec2{(float)n.iparam[0] * f[0], (float)n.iparam[1] * f[1]});
        return bb;
    }
    case GSDF_N_OFFSET2D: {  // :421
        Box2 bb = Bounds2(ch(0));
        if (f[0] > 0) return bb;
        bb.max = {bb.max.x + -f[0], bb.max.y + -f[0]};
        bb.min = {bb.min.x + f[0], bb.min.y + f[0]};
        return bb;
    }
    case GSDF_N_TRANSLATE2D: return Bounds2(ch(0)).addVec({f[0], f[1]});  // :466
    case GSDF_N_ROTATE2D: {  // :514
        Box2 bb = Bounds2(ch(0));
        Mat2 t{f[4], f[5], f[6], f[7]};
        Vec2 verts[4];
        bb.vertices(verts);
        Vec2 v1 = mulMatVec(t, verts[0]);
        bb.max = v1; bb.min = v1;
        for (int i = 1; i < 4; i++) { Vec2 v = mulMatVec(t, verts[i]); bb.max = maxElem(bb.max, v); bb.min = minElem(bb.min, v); }
        return bb;
    }
    case GSDF_N_SYMMETRY2D: {  // :567
        Box2 b = Bounds2(ch(0));
        if (n.iparam[0] & 1) b.min.x = m32::minf(b.min.x, -b.max.x);
        if (n.iparam[0] & 2) b.min.y = m32::minf(b.min.y, -b.max.y);
        return b;
    }
    case GSDF_N_ANNULUS2D: {  // :621
        Box2 bb = Bounds2(ch(0));
        return {{bb.min.x - f[0], bb.min.y - f[0]}, {bb.max.x + f[0], bb.max.y + f[0]}};
    }
    case GSDF_N_CIRCARRAY2D: {  // :674
        Box2 bb = Bounds2(ch(0));
        Vec2 verts[4];
        bb.vertices(verts);
        Mat2 m = rotationMat2(m32::kTwoPiF / (float)n.iparam[1]);
        for (int i = 0; i < n.iparam[0] - 1; i++)
            for (int v = 0; v < 4; v++) { verts[v] = mulMatVec(m, verts[v]); bb = bb.includePoint(verts[v]); }
        return bb;
    }
    case GSDF_N_SCALE2D: return Bounds2(ch(0)).scaleOrigin({f[0], f[0]});  // :728
    case GSDF_N_BOUNDS2: return {{f[0], f[1]}, {f[2], f[3]}};              // glbuild.go:1117
    case GSDF_N_TRANSLATEMULTI2D: {  // :784
        Box2 bb{}, elem = Bounds2(ch(0));
        for (int i = 0; i + 1 < n.aux_cnt; i += 2) bb = bb.unionWith(elem.addVec({a[i], a[i + 1]}));
        return bb;
    }
    case GSDF_N_ELONGATE2D: {  // :832
        Box2 b = Bounds2(ch(0));
        b.max = maxElem(b.max, Vec2{});
        b.max = add(b.max, scale(0.5f, Vec2{f[0], f[1]}));
        b.min = scale(-1, b.max);
        return b;
    }
    }
    return Box2{};
}

}  // namespace gsdfhost
