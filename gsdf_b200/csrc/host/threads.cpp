// threads.cpp -- mirror of forge/threads + the BASELINE example scenes; see threads.h.
// Go's untyped-constant arithmetic is exact and is rounded to float32 only where it meets a float32 variable; the
// double-precision constant expressions below are rounded at the same places.
#include "threads.h"

namespace gsdfhost {
namespace threads {

namespace {
constexpr double kSqrt3 = 1.7320508075688772935274463415058723669428052538103806280558069794;  // iso.go:14
constexpr double kCosd30 = kSqrt3 / 2;                                                           // iso.go:11
constexpr double kSind30 = 0.5;
}  // namespace

float Parameters::HexRadius() const { return HexF2F / (float)(2.0 * kCosd30); }          // threads.go:43-45
float Parameters::HexHeight() const { return 2.0f * HexRadius() * (float)(5.0 / 12.0); }  // threads.go:48-50

// threads.go:225-251
float metricf2f(float radius) {
    static const float table[] = {1.75f, 2, 3.2f, 4, 5, 6, 7, 8, 10, 13, 17, 19, 24, 30, 36, 46, 55, 65, 75, 85, 95};
    float est;
    if (radius < (float)(1.2 / 2)) est = 3.2f * radius;
    else if (radius < (float)(3.8 / 2)) est = 4.5f * radius;
    else if (radius < (float)(4.2 / 2)) est = 4.f * radius;
    else est = 3.5f * radius;
    if (m32::absf(radius - (float)(56. / 2)) < 1) est = 86;
    for (int i = (int)(sizeof(table) / sizeof(table[0])) - 1; i >= 0; i--)
        if (est - 1e-2f > table[i]) return table[i];
    return table[0];
}

Threader Threader::ISO(float D, float P, bool ext) {
    Threader t;
    t.kind = Kind::ISO; t.D = D; t.P = P; t.Ext = ext;
    return t;
}

Threader Threader::Basic(Kind kind, float D, float P) {
    Threader t;
    t.kind = kind; t.D = D; t.P = P;
    return t;
}

Threader Threader::UTS(float D, float TPI, bool ext) {
    Threader t;
    t.kind = Kind::UTS; t.D = D; t.TPI = TPI; t.Ext = ext;
    return t;
}

bool Threader::NPTFromNominal(float nominal, Threader &out) {
    struct Spec { float N, D, tpi, ftof; };
    static const Spec tbl[] = {  // npt.go:44-57
        {(float)(1.0 / 8.0), 0.405f, 27, (float)(11.2 / 25.4)},   {(float)(1.0 / 4.0), 0.540f, 18, (float)(15.7 / 25.4)},
        {(float)(3.0 / 8.0), 0.675f, 18, (float)(17.5 / 25.4)},   {(float)(1.0 / 2.0), 0.840f, 14, (float)(22.4 / 25.4)},
        {(float)(3.0 / 4.0), 1.050f, 14, (float)(26.9 / 25.4)},   {1.0f, 1.315f, 11.5f, (float)(35.1 / 25.4)},
        {(float)(1 + 1.0 / 4.0), 1.660f, 11.5f, (float)(44.5 / 25.4)}, {(float)(1 + 1.0 / 2.0), 1.900f, 11.5f, (float)(50.8 / 25.4)},
        {2, 2.375f, 11.5f, (float)(63.5 / 25.4)},                 {(float)(2 + 1.0 / 2.0), 2.875f, 8, (float)(76.2 / 25.4)},
        {3, 3.500f, 8, (float)(88.9 / 25.4)},                     {4, 4.500f, 8, (float)(117.3 / 25.4)},
    };
    const float lookupTol = (float)(1. / 32.);
    for (const Spec &a : tbl)
        if (m32::absf(a.N - nominal) < lookupTol) {
            out = Threader{};
            out.kind = Kind::NPT; out.D = a.D; out.F2F = a.ftof; out.TPI = a.tpi;
            return true;
        }
    return false;
}

Parameters Threader::ThreadParams() const {
    Parameters p;
    float d = D, pitch = P;
    if (kind == Kind::NPT || kind == Kind::UTS) pitch = 1.0f / TPI;  // npt.go:24, uts.go:20
    if (kind == Kind::Knurl) { d = KRadius * 2; pitch = KPitch; }  // knurl.go:46
    float radius = d / 2;                                 // threads.go:213-223 (basic.ThreadParams)
    p.Name = "basic"; p.Radius = radius; p.Pitch = pitch; p.Starts = 1; p.Taper = 0; p.HexF2F = metricf2f(radius);
    if (kind == Kind::NPT) {
        p.Name = "NPT";
        p.Taper = m32::atan((float)(1.0 / 32.0));  // npt.go:26
        if (F2F > 0) p.HexF2F = F2F;
    }
    if (kind == Kind::Knurl) p.Starts = kstarts;
    return p;
}

NodeId Threader::Thread(Builder &bld, std::string &err) const {
    PolygonBuilder poly;
    if (kind == Kind::Knurl) {  // knurl.go:28-42
        poly.addXY(KPitch / 2, 0);
        poly.addXY(KPitch / 2, KRadius);
        poly.addXY(0, KRadius + KHeight);
        poly.addXY(-KPitch / 2, KRadius);
        poly.addXY(-KPitch / 2, 0);
    } else if (kind == Kind::Acme) {  // acme.go:22-45
        float radius = D / 2;
        float h = radius - 0.5f * P;
        float theta = (float)(29.0 / 2.0) * (float)m32::kPi / 180.0f;
        float delta = 0.25f * P * m32::tan(theta);
        float xOfs0 = 0.25f * P - delta, xOfs1 = 0.25f * P + delta;
        poly.addXY(radius, 0);
        poly.addXY(radius, h);
        poly.addXY(xOfs1, h);
        poly.addXY(xOfs0, radius);
        poly.addXY(-xOfs0, radius);
        poly.addXY(-xOfs1, h);
        poly.addXY(-radius, h);
        poly.addXY(-radius, 0);
    } else if (kind == Kind::ANSIButtress || kind == Kind::PlasticButtress) {
        // ansibuttress.go:22-46 (tangents through math32.Tan) / plasticbuttress.go:26-56 (tangents as constants, more rounding)
        const bool plastic = kind == Kind::PlasticButtress;
        float radius = D / 2, p = P;
        float t0 = plastic ? 1.0f : m32::tan(45.0f * (float)m32::kPi / 180);
        float t1 = plastic ? (float)0.1227845609029046 : m32::tan(7.0f * (float)m32::kPi / 180);
        float sum = plastic ? (float)(1.0 + 0.1227845609029046) : t0 + t1;  // untyped constant sum is exact in Go
        float h0 = p / sum;
        float h1 = ((float)(0.6 / 2.0) * p) + (0.5f * h0);
        float hp = p / 2.0f;
        poly.addXY(p, 0);
        poly.addXY(p, radius);
        if (plastic) {
            poly.addXY(hp - ((h0 - h1) * t1), radius).smooth(0.05f * p, 5);
            poly.addXY(t0 * h0 - hp, radius - h1).smooth(0.15f * p, 5);
            poly.addXY((h0 - h1) * t0 - hp, radius).smooth(0.15f * p, 5);
        } else {
            poly.addXY(hp - ((h0 - h1) * t1), radius);
            poly.addXY(t0 * h0 - hp, radius - h1).smooth(0.0714f * p, 5);
            poly.addXY((h0 - h1) * t0 - hp, radius);
        }
        poly.addXY(-p, radius);
        poly.addXY(-p, 0);
    } else {  // iso.go:37-77 (NPT: ISO{D, 1/TPI} with Ext=false, npt.go:34-36; UTS: ISO{D, 1/TPI, Ext}, uts.go:25-27)
        float d = D, p = (kind == Kind::NPT || kind == Kind::UTS) ? 1.0f / TPI : P;
        bool ext = (kind == Kind::NPT) ? false : Ext;
        float radius = d / 2;
        const double tanTheta = kSind30 / kCosd30;
        float h = p / (float)(2.0 * tanTheta);
        float rMajor = radius;
        float r0 = rMajor - (float)(7.0 / 8.0) * h;
        if (ext) {
            float rRoot = (p / 8.0f) / (float)kCosd30;
            float xOfs = (float)(1.0 / 16.0) * p;
            poly.addXY(p, 0);
            poly.addXY(p, r0 + h);
            poly.addXY(p / 2.0f, r0).smooth(rRoot, 5);
            poly.addXY(xOfs, rMajor);
            poly.addXY(-xOfs, rMajor);
            poly.addXY(-p / 2.0f, r0).smooth(rRoot, 5);
            poly.addXY(-p, r0 + h);
            poly.addXY(-p, 0);
        } else {
            float rMinor = r0 + (float)(1.0 / 4.0) * h;
            float rCrest = (p / 16.0f) / (float)kCosd30;
            float xOfs = (float)(1.0 / 8.0) * p;
            poly.addXY(p, 0);
            poly.addXY(p, rMinor);
            poly.addXY(p / 2 - xOfs, rMinor);
            poly.addXY(0, r0 + h).smooth(rCrest, 5);
            poly.addXY(-p / 2 + xOfs, rMinor);
            poly.addXY(-p, rMinor);
            poly.addXY(-p, 0);
        }
    }
    std::vector<Vec2> verts;
    if (!poly.appendVecs(verts, err)) return -1;
    return bld.NewPolygon(verts);
}

NodeId Screw(Builder &bld, float length, const Threader &t, std::string &err) {  // threads.go:76-96
    if (length <= 0) { err = "need greater than zero length"; return -1; }
    NodeId tsdf = t.Thread(bld, err);
    if (tsdf < 0) return -1;
    Parameters p = t.ThreadParams();
    return bld.NewScrew(tsdf, p.Pitch, -p.Pitch * (float)p.Starts, length, p.Taper);
}

NodeId HexHead(Builder &bld, float radius, float height, bool roundNeg, bool roundPos, std::string &err) {  // hexhead.go:15-47
    float cornerRound = radius * 0.08f;
    PolygonBuilder poly;
    poly.nagon(6, radius - cornerRound);
    std::vector<Vec2> verts;
    if (!poly.appendVecs(verts, err)) return -1;
    NodeId hex2d = bld.NewPolygon(verts);
    hex2d = bld.Offset2D(hex2d, -cornerRound);
    NodeId hex3d = bld.Extrude(hex2d, height);
    if (roundPos || roundNeg) {
        float topRound = radius * 1.6f;
        float d = radius * (float)kCosd30;
        NodeId sphere = bld.NewSphere(topRound);
        float zOfs = m32::sqrt(topRound * topRound - d * d) - height / 2;
        if (roundNeg) hex3d = bld.Intersection(hex3d, bld.Translate(sphere, 0, 0, -zOfs));
        if (roundPos) hex3d = bld.Intersection(hex3d, bld.Translate(sphere, 0, 0, zOfs));
    }
    return hex3d;
}

NodeId Knurl(Builder &bld, Threader k, std::string &err) {  // knurl.go:51-81
    if (k.KLength <= 0) { err = "zero or negative Knurl length"; return -1; }
    if (k.KRadius <= 0) { err = "zero or negative Knurl radius"; return -1; }
    if (k.KPitch <= 0) { err = "zero or negative Knurl pitch"; return -1; }
    if (k.KHeight <= 0) { err = "zero or negative Knurl height"; return -1; }
    if (k.KTheta < 0) { err = "zero Knurl helix angle"; return -1; }
    if (k.KTheta >= (float)(m32::kPi / 2)) { err = "too large Knurl helix angle"; return -1; }
    k.kind = Kind::Knurl;
    k.kstarts = (int)(m32::kTwoPiF * k.KRadius * m32::tan(k.KTheta) / k.KPitch);
    NodeId k0 = Screw(bld, k.KLength, k, err);
    if (k0 < 0) return -1;
    k.kstarts *= -1;
    NodeId k1 = Screw(bld, k.KLength, k, err);
    if (k1 < 0) return -1;
    return bld.Intersection(k0, k1);
}

NodeId KnurledHead(Builder &bld, float radius, float height, float pitch, std::string &err) {  // knurl.go:84-101
    float cylinderRound = radius * 0.05f;
    float knurlLength = pitch * floorf((height - cylinderRound) / pitch);
    Threader k;
    k.kind = Kind::Knurl;
    k.KLength = knurlLength; k.KRadius = radius; k.KPitch = pitch; k.KHeight = pitch * 0.3f;
    k.KTheta = (float)(45.0 * m32::kPi / 180);
    NodeId knurl = Knurl(bld, k, err);
    if (knurl < 0) return -1;
    NodeId cyl = bld.NewCylinder(radius, height, cylinderRound);
    return bld.Union({cyl, knurl});
}

NodeId Nut(Builder &bld, const Threader &t, NutStyle style, float tolerance, std::string &err) {  // nut.go:41-80
    if (tolerance < 0) { err = "tolerance < 0"; return -1; }
    Parameters params = t.ThreadParams();
    float nr = params.HexRadius(), nh = params.HexHeight();
    if (nr <= 0 || nh <= 0) { err = "bad hex nut dimensions"; return -1; }
    NodeId nut = -1;
    switch (style) {
    case NutHex: nut = HexHead(bld, nr, nh, true, true, err); break;
    case NutKnurl: nut = KnurledHead(bld, nr, nh, nr * 0.25f, err); break;
    case NutCircular: nut = bld.NewCylinder(nr * 1.1f, nh, 0); break;
    default: err = "passed argument NutStyle not defined for Nut"; return -1;
    }
    if (nut < 0) return -1;
    NodeId thread = Screw(bld, nh * (float)(1 + 1e-2), t, err);
    if (thread < 0) return -1;
    return bld.Difference(nut, thread);
}

NodeId Bolt(Builder &bld, const Threader &t, NutStyle style, float tolerance, float totalLength, float shankLength,
            std::string &err) {  // bolt.go:21-75
    if (totalLength < 0) { err = "total length < 0"; return -1; }
    if (shankLength >= totalLength) { err = "shank length must be less than total length"; return -1; }
    if (shankLength <= 0) { err = "shank length <= 0"; return -1; }
    if (tolerance < 0) { err = "tolerance < 0"; return -1; }
    Parameters param = t.ThreadParams();
    float hr = param.HexRadius(), hh = param.HexHeight();
    if (hr <= 0 || hh <= 0) { err = "bad hex head dimension"; return -1; }
    NodeId head = -1;
    switch (style) {
    case NutHex: head = HexHead(bld, hr, hh, false, true, err); break;
    case NutKnurl: head = KnurledHead(bld, hr, hh, hr * 0.25f, err); break;
    default: err = "unknown style for bolt"; return -1;
    }
    if (head < 0) return -1;
    float screwLen = totalLength - shankLength;
    NodeId screw = Screw(bld, screwLen, t, err);
    if (screw < 0) return -1;
    NodeId shank = bld.NewCylinder(param.Radius, shankLength, hh * 0.08f);
    float shankOff = shankLength / 2 + hh / 2;
    shank = bld.Translate(shank, 0, 0, shankOff);
    screw = bld.Translate(screw, 0, 0, shankOff + screwLen / 2);
    return bld.Union({screw, bld.SmoothUnion(hh * 0.12f, shank, head)});
}

}  // namespace threads

namespace scenes {

NodeId NptFlange(Builder &bld, std::string &err) {  // examples/npt-flange/flange.go:23-59
    const double tlen = 18. / 25.4, internalDiameter = 1.5 / 2., flangeH = 7. / 25.4, flangeD = 60. / 25.4;
    threads::Threader npt;
    if (!threads::Threader::NPTFromNominal((float)(1.0 / 2.0), npt)) { err = "nominal measurement not found"; return -1; }
    NodeId pipe = threads::Nut(bld, npt, threads::NutCircular, 0, err);
    if (pipe < 0) return -1;
    NodeId flange = bld.NewCylinder((float)(flangeD / 2), (float)flangeH, (float)(flangeH / 8));
    flange = bld.Translate(flange, 0, 0, (float)(-tlen / 2));
    NodeId u = bld.SmoothUnion(0.2f, pipe, flange);
    NodeId hole = bld.NewCylinder((float)(internalDiameter / 2), (float)(4 * flangeH), 0);
    u = bld.Difference(u, hole);
    u = bld.Scale(u, 25.4f);
    err = bld.Err();
    return u;
}

NodeId Bolt(Builder &bld, std::string &err) {  // examples/bolt/main.go:26-41
    const float L = 8, shank = 3;
    threads::Threader th = threads::Threader::ISO(3, 0.5f, true);
    NodeId m3 = threads::Bolt(bld, th, threads::NutHex, 0, L + shank, shank, err);
    if (m3 < 0) return -1;
    m3 = bld.Rotate(m3, (float)(2.5 * m32::kPi / 2), Vec3{1, 0, 0.1f});
    err = bld.Err();
    return m3;
}

NodeId KnurledCylinder(Builder &bld, float diameter, std::string &err) {  // knurled-cyl.go:57-107
    float r = diameter / 2;
    float length = 5 * r, holeDiam = r, knurlSide = r;
    const float smoothRatio = 0.1f, twistK = 0.75f, knurlOffsetR = 1.6f;
    const int knurlN = 24;
    float sk = smoothRatio * r;
    NodeId obj = bld.NewCylinder(r, length, smoothRatio * r);
    NodeId knurlBox = bld.NewBox(knurlSide, knurlSide, length * 0.8f, 0);
    knurlBox = bld.Rotate(knurlBox, (float)(m32::kPi / 4), Vec3{0, 0, 1});
    knurlBox = bld.Translate(knurlBox, knurlOffsetR * r, 0, 0);
    knurlBox = bld.CircularArray(knurlBox, knurlN, knurlN);
    NodeId knurl = bld.Union({bld.Twist(knurlBox, twistK / r), bld.Twist(knurlBox, -twistK / r)});
    obj = bld.SmoothDifference(sk, obj, knurl);
    obj = bld.SmoothDifference(sk, obj, bld.NewCylinder(holeDiam / 2, length + 2 * r, 0));
    NodeId vent = bld.NewCylinder(0.25f * r, 3 * r, 0);
    vent = bld.Rotate(vent, (float)(m32::kPi / 2), Vec3{0, 1, 0});
    obj = bld.SmoothDifference(sk, obj, bld.Translate(vent, 0, 0, -length / 2));
    obj = bld.SmoothDifference(sk, obj, bld.Translate(vent, 0, 0, length / 2));
    err = bld.Err();
    return obj;
}

// examples/fibonacci-showerhead/showerhead.go:31-92 (the PNG side output of :55-56 is not part of the shape)
NodeId FibonacciShowerhead(Builder &bld, std::string &err) {
    const double threadExtDiameter = 65., threadedLength = 5., threadTurns = 3., threadPitch = threadedLength / threadTurns;
    const double showerheadBaseThick = 2.5, showerheadWall = 4., threadheight = 5.;
    const threads::Threader showerThread = threads::Threader::Basic(threads::Kind::PlasticButtress, (float)threadExtDiameter, (float)threadPitch);
    NodeId knurled = threads::KnurledHead(bld, (float)(threadExtDiameter / 2 + showerheadWall), (float)threadheight, 1, err);
    if (knurled < 0) return -1;
    NodeId thr = threads::Screw(bld, (float)(threadheight + .5), showerThread, err);
    if (thr < 0) return -1;
    NodeId object = bld.Difference(knurled, thr);
    NodeId base = bld.NewCylinder((float)(threadExtDiameter / 2 + showerheadWall), (float)showerheadBaseThick, 0);
    base = bld.Translate(base, 0, 0, (float)-(threadedLength / 2 + showerheadBaseThick / 2 - 1));
    NodeId hole = bld.NewCylinder(0.8f, (float)(showerheadBaseThick * 10), 0);
    NodeId holes = hole;
    for (int i = 0; i < 130; i++) {  // fibonacci(i), showerhead.go:137-146
        const float nf = (float)i;
        const float a = nf * 137.3f / 360 * (float)m32::kPi;
        const float r = 2.6f * m32::sqrt(nf);
        float sa, ca;
        m32::sincos(a, sa, ca);
        holes = bld.Union({holes, bld.Translate(hole, r * ca, r * sa, 0)});
    }
    base = bld.Difference(base, holes);
    object = bld.Union({object, base});
    const std::string berr = bld.Err();
    if (!berr.empty()) { err = berr; return -1; }
    return object;
}

}  // namespace scenes
}  // namespace gsdfhost
