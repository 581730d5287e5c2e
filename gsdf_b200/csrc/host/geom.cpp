// geom.cpp -- see geom.h. Restated soypat/geometry helpers (UNPINNED against the Go module).
#include "geom.h"

namespace gsdfhost {

float determinant(const Mat4 &a) {
    const float(*m)[4] = a.x;
    return m[0][0] * m[1][1] * m[2][2] * m[3][3] - m[0][0] * m[1][1] * m[2][3] * m[3][2] + m[0][0] * m[1][2] * m[2][3] * m[3][1] -
           m[0][0] * m[1][2] * m[2][1] * m[3][3] + m[0][0] * m[1][3] * m[2][1] * m[3][2] - m[0][0] * m[1][3] * m[2][2] * m[3][1] -
           m[0][1] * m[1][2] * m[2][3] * m[3][0] + m[0][1] * m[1][2] * m[2][0] * m[3][3] - m[0][1] * m[1][3] * m[2][0] * m[3][2] +
           m[0][1] * m[1][3] * m[2][2] * m[3][0] - m[0][1] * m[1][0] * m[2][2] * m[3][3] + m[0][1] * m[1][0] * m[2][3] * m[3][2] +
           m[0][2] * m[1][3] * m[2][0] * m[3][1] - m[0][2] * m[1][3] * m[2][1] * m[3][0] + m[0][2] * m[1][0] * m[2][1] * m[3][3] -
           m[0][2] * m[1][0] * m[2][3] * m[3][1] + m[0][2] * m[1][1] * m[2][3] * m[3][0] - m[0][2] * m[1][1] * m[2][0] * m[3][3] -
           m[0][3] * m[1][0] * m[2][1] * m[3][2] + m[0][3] * m[1][0] * m[2][2] * m[3][1] - m[0][3] * m[1][1] * m[2][2] * m[3][0] +
           m[0][3] * m[1][1] * m[2][0] * m[3][2] - m[0][3] * m[1][2] * m[2][0] * m[3][1] + m[0][3] * m[1][2] * m[2][1] * m[3][0];
}

// General 4x4 inverse by cofactors (adjugate / determinant), as ms3.Mat4.Inverse does.
Mat4 inverse(const Mat4 &a) {
    const float(*m)[4] = a.x;
    Mat4 r;
    float d = determinant(a);
    float id = 1 / d;
    r.x[0][0] = (m[1][2] * m[2][3] * m[3][1] - m[1][3] * m[2][2] * m[3][1] + m[1][3] * m[2][1] * m[3][2] - m[1][1] * m[2][3] * m[3][2] - m[1][2] * m[2][1] * m[3][3] + m[1][1] * m[2][2] * m[3][3]) * id;
    r.x[0][1] = (m[0][3] * m[2][2] * m[3][1] - m[0][2] * m[2][3] * m[3][1] - m[0][3] * m[2][1] * m[3][2] + m[0][1] * m[2][3] * m[3][2] + m[0][2] * m[2][1] * m[3][3] - m[0][1] * m[2][2] * m[3][3]) * id;
    r.x[0][2] = (m[0][2] * m[1][3] * m[3][1] - m[0][3] * m[1][2] * m[3][1] + m[0][3] * m[1][1] * m[3][2] - m[0][1] * m[1][3] * m[3][2] - m[0][2] * m[1][1] * m[3][3] + m[0][1] * m[1][2] * m[3][3]) * id;
    r.x[0][3] = (m[0][3] * m[1][2] * m[2][1] - m[0][2] * m[1][3] * m[2][1] - m[0][3] * m[1][1] * m[2][2] + m[0][1] * m[1][3] * m[2][2] + m[0][2] * m[1][1] * m[2][3] - m[0][1] * m[1][2] * m[2][3]) * id;
    r.x[1][0] = (m[1][3] * m[2][2] * m[3][0] - m[1][2] * m[2][3] * m[3][0] - m[1][3] * m[2][0] * m[3][2] + m[1][0] * m[2][3] * m[3][2] + m[1][2] * m[2][0] * m[3][3] - m[1][0] * m[2][2] * m[3][3]) * id;
    r.x[1][1] = (m[0][2] * m[2][3] * m[3][0] - m[0][3] * m[2][2] * m[3][0] + m[0][3] * m[2][0] * m[3][2] - m[0][0] * m[2][3] * m[3][2] - m[0][2] * m[2][0] * m[3][3] + m[0][0] * m[2][2] * m[3][3]) * id;
    r.x[1][2] = (m[0][3] * m[1][2] * m[3][0] - m[0][2] * m[1][3] * m[3][0] - m[0][3] * m[1][0] * m[3][2] + m[0][0] * m[1][3] * m[3][2] + m[0][2] * m[1][0] * m[3][3] - m[0][0] * m[1][2] * m[3][3]) * id;
    r.x[1][3] = (m[0][2] * m[1][3] * m[2][0] - m[0][3] * m[1][2] * m[2][0] + m[0][3] * m[1][0] * m[2][2] - m[0][0] * m[1][3] * m[2][2] - m[0][2] * m[1][0] * m[2][3] + m[0][0] * m[1][2] * m[2][3]) * id;
    r.x[2][0] = (m[1][1] * m[2][3] * m[3][0] - m[1][3] * m[2][1] * m[3][0] + m[1][3] * m[2][0] * m[3][1] - m[1][0] * m[2][3] * m[3][1] - m[1][1] * m[2][0] * m[3][3] + m[1][0] * m[2][1] * m[3][3]) * id;
    r.x[2][1] = (m[0][3] * m[2][1] * m[3][0] - m[0][1] * m[2][3] * m[3][0] - m[0][3] * m[2][0] * m[3][1] + m[0][0] * m[2][3] * m[3][1] + m[0][1] * m[2][0] * m[3][3] - m[0][0] * m[2][1] * m[3][3]) * id;
    r.x[2][2] = (m[0][1] * m[1][3] * m[3][0] - m[0][3] * m[1][1] * m[3][0] + m[0][3] * m[1][0] * m[3][1] - m[0][0] * m[1][3] * m[3][1] - m[0][1] * m[1][0] * m[3][3] + m[0][0] * m[1][1] * m[3][3]) * id;
    r.x[2][3] = (m[0][3] * m[1][1] * m[2][0] - m[0][1] * m[1][3] * m[2][0] - m[0][3] * m[1][0] * m[2][1] + m[0][0] * m[1][3] * m[2][1] + m[0][1] * m[1][0] * m[2][3] - m[0][0] * m[1][1] * m[2][3]) * id;
    r.x[3][0] = (m[1][2] * m[2][1] * m[3][0] - m[1][1] * m[2][2] * m[3][0] - m[1][2] * m[2][0] * m[3][1] + m[1][0] * m[2][2] * m[3][1] + m[1][1] * m[2][0] * m[3][2] - m[1][0] * m[2][1] * m[3][2]) * id;
    r.x[3][1] = (m[0][1] * m[2][2] * m[3][0] - m[0][2] * m[2][1] * m[3][0] + m[0][2] * m[2][0] * m[3][1] - m[0][0] * m[2][2] * m[3][1] - m[0][1] * m[2][0] * m[3][2] + m[0][0] * m[2][1] * m[3][2]) * id;
    r.x[3][2] = (m[0][2] * m[1][1] * m[3][0] - m[0][1] * m[1][2] * m[3][0] - m[0][2] * m[1][0] * m[3][1] + m[0][0] * m[1][2] * m[3][1] + m[0][1] * m[1][0] * m[3][2] - m[0][0] * m[1][1] * m[3][2]) * id;
    r.x[3][3] = (m[0][1] * m[1][2] * m[2][0] - m[0][2] * m[1][1] * m[2][0] + m[0][2] * m[1][0] * m[2][1] - m[0][0] * m[1][2] * m[2][1] - m[0][1] * m[1][0] * m[2][2] + m[0][0] * m[1][1] * m[2][2]) * id;
    return r;
}

Box3 mulBox(const Mat4 &a, const Box3 &box) {
    Vec3 r{a.x[0][0], a.x[1][0], a.x[2][0]};
    Vec3 u{a.x[0][1], a.x[1][1], a.x[2][1]};
    Vec3 b{a.x[0][2], a.x[1][2], a.x[2][2]};
    Vec3 t{a.x[0][3], a.x[1][3], a.x[2][3]};
    Vec3 xa = scale(box.min.x, r), xb = scale(box.max.x, r);
    Vec3 ya = scale(box.min.y, u), yb = scale(box.max.y, u);
    Vec3 za = scale(box.min.z, b), zb = scale(box.max.z, b);
    Vec3 xmin = minElem(xa, xb), xmax = maxElem(xa, xb);
    Vec3 ymin = minElem(ya, yb), ymax = maxElem(ya, yb);
    Vec3 zmin = minElem(za, zb), zmax = maxElem(za, zb);
    return {add(add(add(xmin, ymin), zmin), t), add(add(add(xmax, ymax), zmax), t)};
}

void PolygonBuilder::nagon(int n, float centerDistance) {
    if (n < 3) return;
    Mat2 m = rotationMat2(m32::kTwoPiF / (float)n);
    Vec2 v{centerDistance, 0};
    for (int i = 0; i < n; i++) {
        addXY(v.x, v.y);
        v = mulMatVec(m, v);
    }
}

bool PolygonBuilder::appendVecs(std::vector<Vec2> &out, std::string &err) const {
    size_t n = verts_.size();
    if (n < 2) { err = "too few vertices"; return false; }
    for (size_t i = 0; i < n; i++) {
        const PV &v = verts_[i];
        if (v.radius == 0 || v.facets <= 0) { out.push_back(v.v); continue; }
        // Replace the vertex by a tangent circular arc of `facets` segments (facets+1 points).
        const PV &vp = verts_[(i + n - 1) % n];
        const PV &vn = verts_[(i + 1) % n];
        Vec2 dp = sub(vp.v, v.v), dn = sub(vn.v, v.v);
        Vec2 v0 = unit(dp), v1 = unit(dn);
        float theta = m32::acos(dot(v0, v1));
        float d1 = v.radius / m32::tan(theta / 2);
        if (d1 > norm(dp) || d1 > norm(dn)) { err = "unable to smooth polygon vertex: radius too large"; return false; }
        Vec2 p0 = add(v.v, scale(d1, v0));
        float d2 = v.radius / m32::sin(theta / 2);
        Vec2 vc = unit(add(v0, v1));
        Vec2 c = add(v.v, scale(d2, vc));
        float sgn = m32::signf(cross(v1, v0));
        float dtheta = sgn * (m32::kPiF - theta) / (float)v.facets;
        Mat2 rm = rotationMat2(dtheta);
        Vec2 rv = sub(p0, c);
        for (int j = 0; j <= v.facets; j++) {
            out.push_back(add(c, rv));
            rv = mulMatVec(rm, rv);
        }
    }
    return true;
}

}  // namespace gsdfhost
