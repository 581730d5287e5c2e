// geom.h -- host-side float32 vector / box / matrix helpers and the polygon builder.
//
// The reference takes these from github.com/soypat/geometry v0.0.0-20251107203642-291c5648d529 (go.mod:11; packages
// ms2, ms3), which is NOT vendored under the reference tree.  They are restated here from that module's published
// behaviour (gonum r2/r3-style vector ops, sdfx-style polygon smoothing / box transforms).  Their outputs (polygon
// vertices, 4x4 inverses, bounding boxes) are INPUTS to both the CUDA kernels and the CPU oracle, so a last-bit
// difference from the Go module moves the scene, not the kernel-vs-oracle comparison ("parity unpinned").
#pragma once
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../math32.cuh"

namespace gsdfhost {

struct Vec2 {
    float x = 0, y = 0;
};
struct Vec3 {
    float x = 0, y = 0, z = 0;
};

inline Vec2 add(Vec2 a, Vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline Vec2 sub(Vec2 a, Vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline Vec2 scale(float s, Vec2 a) { return {s * a.x, s * a.y}; }
inline float dot(Vec2 a, Vec2 b) { return a.x * b.x + a.y * b.y; }
inline float cross(Vec2 a, Vec2 b) { return a.x * b.y - a.y * b.x; }
inline float norm(Vec2 a) { return m32::hypot32(a.x, a.y); }
inline Vec2 unit(Vec2 a) { return scale(1 / norm(a), a); }
inline Vec2 minElem(Vec2 a, Vec2 b) { return {m32::minf(a.x, b.x), m32::minf(a.y, b.y)}; }
inline Vec2 maxElem(Vec2 a, Vec2 b) { return {m32::maxf(a.x, b.x), m32::maxf(a.y, b.y)}; }
inline bool operator==(Vec2 a, Vec2 b) { return a.x == b.x && a.y == b.y; }

inline Vec3 add(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 sub(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 scale(float s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 mulElem(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline Vec3 addScalar(float s, Vec3 a) { return {a.x + s, a.y + s, a.z + s}; }
inline float norm(Vec3 a) { return m32::norm3(a.x, a.y, a.z); }
inline Vec3 unit(Vec3 a) { return scale(1 / norm(a), a); }
inline Vec3 minElem(Vec3 a, Vec3 b) { return {m32::minf(a.x, b.x), m32::minf(a.y, b.y), m32::minf(a.z, b.z)}; }
inline Vec3 maxElem(Vec3 a, Vec3 b) { return {m32::maxf(a.x, b.x), m32::maxf(a.y, b.y), m32::maxf(a.z, b.z)}; }
inline float maxComp(Vec3 a) { return m32::maxf(a.x, m32::maxf(a.y, a.z)); }
inline float minComp(Vec3 a) { return m32::minf(a.x, m32::minf(a.y, a.z)); }

// ms2.Mat2 {x00 x01 / x10 x11}; RotationMat2(a) = {c,-s / s,c}
struct Mat2 {
    float x00 = 1, x01 = 0, x10 = 0, x11 = 1;
};
inline Mat2 rotationMat2(float a) {
    float s, c;
    m32::sincos(a, s, c);
    return {c, -s, s, c};
}
inline Vec2 mulMatVec(const Mat2 &m, Vec2 v) { return {m.x00 * v.x + m.x01 * v.y, m.x10 * v.x + m.x11 * v.y}; }
inline float determinant(const Mat2 &m) { return m.x00 * m.x11 - m.x01 * m.x10; }
inline Mat2 inverse(const Mat2 &m) {
    float d = 1 / determinant(m);
    return {m.x11 * d, -m.x01 * d, -m.x10 * d, m.x00 * d};
}

struct Box2 {
    Vec2 min, max;
    Vec2 size() const { return sub(max, min); }
    bool empty() const { return min.x >= max.x || min.y >= max.y; }
    Box2 unionWith(const Box2 &b) const {
        if (empty()) return b;
        if (b.empty()) return *this;
        return {minElem(min, b.min), maxElem(max, b.max)};
    }
    Box2 intersect(const Box2 &b) const {
        Box2 r{maxElem(min, b.min), minElem(max, b.max)};
        if (r.empty()) return Box2{};
        return r;
    }
    Box2 addVec(Vec2 v) const { return {add(min, v), add(max, v)}; }
    Box2 scaleOrigin(Vec2 s) const { return Box2{{min.x * s.x, min.y * s.y}, {max.x * s.x, max.y * s.y}}.canon(); }
    Box2 canon() const { return {minElem(min, max), maxElem(min, max)}; }
    Box2 includePoint(Vec2 p) const { return {minElem(min, p), maxElem(max, p)}; }
    void vertices(Vec2 out[4]) const {
        out[0] = {min.x, min.y}; out[1] = {max.x, min.y}; out[2] = {max.x, max.y}; out[3] = {min.x, max.y};
    }
};

struct Box3 {
    Vec3 min, max;
    Vec3 size() const { return sub(max, min); }
    Vec3 center() const { return add(min, scale(0.5f, size())); }
    float diagonal() const { return norm(size()); }
    bool empty() const { return min.x >= max.x || min.y >= max.y || min.z >= max.z; }
    Box3 unionWith(const Box3 &b) const {
        if (empty()) return b;
        if (b.empty()) return *this;
        return {minElem(min, b.min), maxElem(max, b.max)};
    }
    Box3 intersect(const Box3 &b) const {
        Box3 r{maxElem(min, b.min), minElem(max, b.max)};
        if (r.empty()) return Box3{};
        return r;
    }
    Box3 addVec(Vec3 v) const { return {add(min, v), add(max, v)}; }
    Box3 scaleOrigin(Vec3 s) const { return Box3{mulElem(min, s), mulElem(max, s)}.canon(); }
    Box3 canon() const { return {minElem(min, max), maxElem(min, max)}; }
    void vertices(Vec3 out[8]) const {
        out[0] = {min.x, min.y, min.z}; out[1] = {max.x, min.y, min.z}; out[2] = {max.x, max.y, min.z};
        out[3] = {min.x, max.y, min.z}; out[4] = {min.x, min.y, max.z}; out[5] = {max.x, min.y, max.z};
        out[6] = {max.x, max.y, max.z}; out[7] = {min.x, max.y, max.z};
    }
};
inline Box3 centeredBox(Vec3 c, Vec3 size) {
    Vec3 h = scale(0.5f, size);
    return {sub(c, h), add(c, h)};
}

// ms3.Mat4, row-major x[r][c]
struct Mat4 {
    float x[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
};
// ms3.RotationMat4(angle, axis): Rodrigues with a normalised axis.
inline Mat4 rotationMat4(float angle, Vec3 axis) {
    axis = unit(axis);
    float s, c;
    m32::sincos(angle, s, c);
    float m = 1 - c;
    Mat4 r;
    r.x[0][0] = m * axis.x * axis.x + c;          r.x[0][1] = m * axis.x * axis.y - axis.z * s; r.x[0][2] = m * axis.z * axis.x + axis.y * s; r.x[0][3] = 0;
    r.x[1][0] = m * axis.x * axis.y + axis.z * s; r.x[1][1] = m * axis.y * axis.y + c;          r.x[1][2] = m * axis.y * axis.z - axis.x * s; r.x[1][3] = 0;
    r.x[2][0] = m * axis.z * axis.x - axis.y * s; r.x[2][1] = m * axis.y * axis.z + axis.x * s; r.x[2][2] = m * axis.z * axis.z + c;          r.x[2][3] = 0;
    r.x[3][0] = 0; r.x[3][1] = 0; r.x[3][2] = 0; r.x[3][3] = 1;
    return r;
}
float determinant(const Mat4 &m);
Mat4 inverse(const Mat4 &m);
// ms3.Mat4.MulBox: transform an AABB and re-fit (Arvo's method).
Box3 mulBox(const Mat4 &a, const Box3 &b);

// ms2.PolygonBuilder: AddXY / Smooth / Chamfer / Nagon / AppendVecs.
class PolygonBuilder {
public:
    PolygonBuilder &addXY(float x, float y) {
        verts_.push_back({{x, y}, 0.f, 0});
        return *this;
    }
    // apply to the vertex added last
    PolygonBuilder &smooth(float radius, int facets) {
        if (!verts_.empty()) { verts_.back().radius = radius; verts_.back().facets = facets; }
        return *this;
    }
    PolygonBuilder &chamfer(float size) { return smooth(size * 1.4142135623730951f, 1); }
    void nagon(int n, float centerDistance);
    // returns false and sets err on failure
    bool appendVecs(std::vector<Vec2> &out, std::string &err) const;

private:
    struct PV {
        Vec2 v;
        float radius;
        int facets;
    };
    std::vector<PV> verts_;
};

}  // namespace gsdfhost
