/*
 * gsdf_oracle.c -- CPU ORACLE (test infrastructure, see gsdf_oracle.h for scope and parity-pinning status).
 *
 * Literal float32 restatement of the reference's CPU evaluators and mesher.  Build with
 *     gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared
 * (Go on amd64 never fuses a*b+c, so contraction must be off for the operation sequences below to round the same).
 * Batched, recursive, scratch-buffer structure mirrors cpu_evaluators.go so it is also a fair CPU baseline.
 */
#include "gsdf_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "mc_tables.inc"

/* ------------------------------------------------------------------------------------------------------------
 * float32 math: chewxy/math32 v1.11.1 (go.mod:8) restated. That module is a float32 port of Go's math package
 * (Cephes-derived); it is not vendored, so last-bit parity with it is UNPINNED.
 * ---------------------------------------------------------------------------------------------------------- */
#define GO_PI 3.14159265358979323846264338327950288419716939937510582097494459

float go_sqrt(float x) { return sqrtf(x); } /* SQRTSS, IEEE correctly rounded */

/* math32.Min / math32.Max: Go semantics (NaN propagates, -0 < +0). gsdf.go:141-143,169-171 */
float go_min(float x, float y) {
    if (isinf(x) && x < 0) return x;
    if (isinf(y) && y < 0) return y;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0 && x == y) return signbit(x) ? x : y;
    return x < y ? x : y;
}
float go_max(float x, float y) {
    if (isinf(x) && x > 0) return x;
    if (isinf(y) && y > 0) return y;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0 && x == y) return signbit(x) ? y : x;
    return x > y ? x : y;
}
static inline float go_abs(float x) { return fabsf(x); }

/* math32.Hypot: p*sqrt(1+(q/p)^2) with p>=q, not sqrt(p*p+q*q). */
float go_hypot(float p, float q) {
    if (isinf(p) || isinf(q)) return INFINITY;
    if (isnan(p) || isnan(q)) return NAN;
    if (p < 0) p = -p;
    if (q < 0) q = -q;
    if (p < q) { float t = p; p = q; q = t; }
    if (p == 0) return 0;
    q = q / p;
    return p * sqrtf(1 + q * q);
}

float go_floor(float x) { return floorf(x); }  /* exact */
float go_round(float x) { return roundf(x); }  /* half away from zero, as Go's Round */

/* math32 atan.go: Cephes rational approximation, float32 arithmetic. */
static float go_xatan(float x) {
    const float P0 = -8.750608600031904122785e-01f, P1 = -1.615753718733365076637e+01f,
                P2 = -7.500855792314704667340e+01f, P3 = -1.228866684490136173410e+02f,
                P4 = -6.485021904942025371773e+01f;
    const float Q0 = +2.485846490142306297962e+01f, Q1 = +1.650270098316988542046e+02f,
                Q2 = +4.328810604912902668951e+02f, Q3 = +4.853903996359136964868e+02f,
                Q4 = +1.945506571482613964425e+02f;
    float z = x * x;
    z = z * ((((P0 * z + P1) * z + P2) * z + P3) * z + P4) / (((((z + Q0) * z + Q1) * z + Q2) * z + Q3) * z + Q4);
    z = x * z + x;
    return z;
}
static float go_satan(float x) {
    const float Morebits = 6.123233995736765886130e-17f;
    const float Tan3pio8 = 2.41421356237309504880f;
    if (x <= 0.66f) return go_xatan(x);
    if (x > Tan3pio8) return (float)(GO_PI / 2) - go_xatan(1 / x) + Morebits;
    return (float)(GO_PI / 4) + go_xatan((x - 1) / (x + 1)) + 0.5f * Morebits;
}
float go_atan(float x) {
    if (x == 0) return x;
    if (x > 0) return go_satan(x);
    return -go_satan(-x);
}
float go_atan2(float y, float x) {
    const float pi = (float)GO_PI;
    if (isnan(y) || isnan(x)) return NAN;
    if (y == 0) {
        if (x >= 0 && !signbit(x)) return copysignf(0, y);
        return copysignf(pi, y);
    }
    if (x == 0) return copysignf((float)(GO_PI / 2), y);
    if (isinf(x)) {
        if (x > 0) return isinf(y) ? copysignf((float)(GO_PI / 4), y) : copysignf(0, y);
        return isinf(y) ? copysignf((float)(3 * GO_PI / 4), y) : copysignf(pi, y);
    }
    if (isinf(y)) return copysignf((float)(GO_PI / 2), y);
    float q = go_atan(y / x);
    if (x < 0) return q <= 0 ? q + pi : q - pi;
    return q;
}

/* math32 sin.go: Cephes sin/cos, 3-part pi/4 reduction, float32 arithmetic. */
static const float go_sinc[6] = {1.58962301576546568060e-10f, -2.50507477628578072866e-8f,
                                 2.75573136213857245213e-6f,  -1.98412698295895385996e-4f,
                                 8.33333333332211858878e-3f,  -1.66666666666666307295e-1f};
static const float go_cosc[6] = {-1.13585365213876817300e-11f, 2.08757008419747316778e-9f,
                                 -2.75573141792967388112e-7f,  2.48015872888517045348e-5f,
                                 -1.38888888888730564116e-3f,  4.16666666666665929218e-2f};
#define GO_PI4A 7.85398125648498535156e-1f
#define GO_PI4B 3.77489470793079817668e-8f
#define GO_PI4C 2.69515142907905952645e-15f
#define GO_M4PI 1.273239544735162542821171882678754627704620361328125f

static inline float go_sin_poly(float z, float zz) {
    return z + z * zz * ((((((go_sinc[0] * zz) + go_sinc[1]) * zz + go_sinc[2]) * zz + go_sinc[3]) * zz + go_sinc[4]) * zz + go_sinc[5]);
}
static inline float go_cos_poly(float zz) {
    return 1.0f - 0.5f * zz + zz * zz * ((((((go_cosc[0] * zz) + go_cosc[1]) * zz + go_cosc[2]) * zz + go_cosc[3]) * zz + go_cosc[4]) * zz + go_cosc[5]);
}
static inline float go_trig_reduce(float x, uint32_t *jout) {
    uint32_t j = (uint32_t)(x * GO_M4PI);
    float y = (float)j;
    if (j & 1) { j++; y += 1; }
    *jout = j & 7;
    return ((x - y * GO_PI4A) - y * GO_PI4B) - y * GO_PI4C;
}
float go_sin(float x) {
    if (x == 0 || isnan(x)) return x;
    if (isinf(x)) return NAN;
    int sign = 0;
    if (x < 0) { x = -x; sign = 1; }
    uint32_t j;
    float z = go_trig_reduce(x, &j);
    if (j > 3) { sign = !sign; j -= 4; }
    float zz = z * z;
    float y = (j == 1 || j == 2) ? go_cos_poly(zz) : go_sin_poly(z, zz);
    return sign ? -y : y;
}
float go_cos(float x) {
    if (isnan(x) || isinf(x)) return NAN;
    int sign = 0;
    x = fabsf(x);
    uint32_t j;
    float z = go_trig_reduce(x, &j);
    if (j > 3) { j -= 4; sign = !sign; }
    if (j > 1) sign = !sign;
    float zz = z * z;
    float y = (j == 1 || j == 2) ? go_sin_poly(z, zz) : go_cos_poly(zz);
    return sign ? -y : y;
}
/* math32 tan.go */
float go_tan(float x) {
    const float P0 = -1.30936939181383777646e4f, P1 = 1.15351664838587416140e6f, P2 = -1.79565251976484877988e7f;
    const float Q1 = 1.36812963470692954678e4f, Q2 = -1.32089234440210967447e6f, Q3 = 2.50083801823357915839e7f,
                Q4 = -5.38695755929454629881e7f;
    if (x == 0 || isnan(x)) return x;
    if (isinf(x)) return NAN;
    int sign = 0;
    if (x < 0) { x = -x; sign = 1; }
    uint32_t j = (uint32_t)(x * GO_M4PI);
    float y = (float)j;
    if (j & 1) { j++; y += 1; }
    float z = ((x - y * GO_PI4A) - y * GO_PI4B) - y * GO_PI4C;
    float zz = z * z;
    if (zz > 1e-14f)
        y = z + z * (zz * (((P0 * zz) + P1) * zz + P2) / ((((zz + Q1) * zz + Q2) * zz + Q3) * zz + Q4));
    else
        y = z;
    if (j & 2) y = -1 / y;
    return sign ? -y : y;
}

/* math32 asin.go: Asin via Sqrt + satan; Acos = Pi/2 - Asin. */
static float go_asin(float x) {
    if (x == 0) return x;
    int sign = 0;
    if (x < 0) { x = -x; sign = 1; }
    if (x > 1) return NAN;
    float temp = sqrtf(1 - x * x);
    if (x > 0.7f) temp = (float)(GO_PI / 2) - go_satan(temp / x);
    else temp = go_satan(x / temp);
    return sign ? -temp : temp;
}
float go_acos(float x) { return (float)(GO_PI / 2) - go_asin(x); }

/* math32.Cbrt restated as FreeBSD's cbrtf (integer seed + two double-precision Newton steps, result rounded once to
 * float32): pure arithmetic, so the CUDA side can run the identical sequence. UNPINNED against math32's own choice. */
float go_cbrt(float x) {
    uint32_t hx;
    memcpy(&hx, &x, 4);
    const uint32_t sign = hx & 0x80000000u;
    hx ^= sign;
    if (hx >= 0x7f800000u) return x + x;
    float t;
    if (hx < 0x00800000u) {
        if (hx == 0) return x;
        uint32_t two24 = 0x4b800000u, high;
        memcpy(&t, &two24, 4);
        t *= x;
        memcpy(&high, &t, 4);
        high = sign | ((high & 0x7fffffffu) / 3 + 642849266u);
        memcpy(&t, &high, 4);
    } else {
        uint32_t w = sign | (hx / 3 + 709958130u);
        memcpy(&t, &w, 4);
    }
    double T = t, r = T * T * T;
    T = T * ((double)x + x + r) / (x + r + r);
    r = T * T * T;
    T = T * ((double)x + x + r) / (x + r + r);
    return (float)T;
}

/* math32 log.go / exp.go (ports of Go's math.Log / math.Exp, FreeBSD e_log.c / e_exp.c), float32 arithmetic. */
float go_log(float x) {
    const float Ln2Hi = 6.93147180369123816490e-01f, Ln2Lo = 1.90821492927058770002e-10f;
    const float L1 = 6.666666666666735130e-01f, L2 = 3.999999999940941908e-01f, L3 = 2.857142874366239149e-01f,
                L4 = 2.222219843214978396e-01f, L5 = 1.818357216161805012e-01f, L6 = 1.531383769920937332e-01f,
                L7 = 1.479819860511658591e-01f;
    if (isnan(x) || (isinf(x) && x > 0)) return x;
    if (x < 0) return NAN;
    if (x == 0) return -INFINITY;
    int ki;
    float f1 = frexpf(x, &ki);
    if (f1 < (float)(1.41421356237309504880168872420969808 / 2)) { f1 *= 2; ki--; }
    float f = f1 - 1;
    float k = (float)ki;
    float s_ = f / (2 + f);
    float s2 = s_ * s_;
    float s4 = s2 * s2;
    float t1 = s2 * (L1 + s4 * (L3 + s4 * (L5 + s4 * L7)));
    float t2 = s4 * (L2 + s4 * (L4 + s4 * L6));
    float R = t1 + t2;
    float hfsq = 0.5f * f * f;
    return k * Ln2Hi - ((hfsq - (s_ * (hfsq + R) + k * Ln2Lo)) - f);
}
float go_exp(float x) {
    const float Ln2Hi = 6.93147180369123816490e-01f, Ln2Lo = 1.90821492927058770002e-10f, Log2e = 1.44269504088896338700e+00f;
    const float P1 = 1.66666666666666657415e-01f, P2 = -2.77777777770155933842e-03f, P3 = 6.61375632143793436117e-05f,
                P4 = -1.65339022054652515390e-06f, P5 = 4.13813679705723846039e-08f;
    if (isnan(x) || (isinf(x) && x > 0)) return x;
    if (isinf(x)) return 0;
    if (x > 88.72283905206835f) return INFINITY;
    if (x < -103.97207708f) return 0;
    int k = 0;
    if (x < 0) k = (int)(Log2e * x - 0.5f);
    else if (x > 0) k = (int)(Log2e * x + 0.5f);
    float hi = x - (float)k * Ln2Hi;
    float lo = (float)k * Ln2Lo;
    float r = hi - lo;
    float t = r * r;
    float c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    float y = 1 - ((lo - (r * c) / (2 - c)) - hi);
    return ldexpf(y, k);
}
/* math32.Pow(x, y) for x >= 0 and 0 < y < 0.5 (the only use on this path: powelem2(1./3, |x|), gsdf.go:182):
 * yi = 0, so the result is Exp(yf*Log(x)) (Go pow.go). */
static float go_pow_frac(float x, float y) {
    if (x == 0) return 0;
    if (x == 1) return 1;
    return go_exp(y * go_log(x));
}

/* gsdf.go:148-167 */
static inline float signf(float a) { return a == 0 ? 0 : copysignf(1, a); }
static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline float mixf(float x, float y, float a) { return x * (1 - a) + y * a; }
/* gsdf.go:16-21 */
#define GSDF_TRIBISECT 0.8660254037844386467637231707529361834714026269051903140279034897
#define GSDF_SQRT3 1.7320508075688772935274463415058723669428052538103806280558069794
#define GSDF_LARGENUM 1e20f

/* soypat/geometry ms3.Norm / ms2.Norm (gonum r3/r2 style, UNPINNED): nested Hypot. */
static inline float norm3(float x, float y, float z) { return go_hypot(x, go_hypot(y, z)); }
static inline float norm2(float x, float y) { return go_hypot(x, y); }

/* ------------------------------------------------------------------------------------------------------------
 * Evaluators. Batched and recursive exactly like cpu_evaluators.go; scratch comes from malloc (VecPool role).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { float x, y, z; } v3;
typedef struct { float x, y; } v2;

static int eval3(const go_tree *t, int id, const v3 *pos, float *dist, size_t n);
static int eval2(const go_tree *t, int id, const v2 *pos, float *dist, size_t n);

static inline const go_node *node_at(const go_tree *t, int id) {
    if (id < 0 || id >= t->nnodes) return NULL;
    return &t->nodes[id];
}
static inline int child_id(const go_tree *t, const go_node *nd, int k) { return t->children[nd->child_off + k]; }

#define SCRATCH(type, name, count)                            \
    type *name = (type *)malloc(sizeof(type) * (count ? count : 1)); \
    if (!name) return -2

static int eval3(const go_tree *t, int id, const v3 *pos, float *dist, size_t n) {
    const go_node *nd = node_at(t, id);
    if (!nd) return -1;
    const float *f = nd->fparam;
    int err = 0;
    switch (nd->kind) {
    case GO_SPHERE: { /* cpu_evaluators.go:20-26 */
        float r = f[0];
        for (size_t i = 0; i < n; i++) dist[i] = norm3(pos[i].x, pos[i].y, pos[i].z) - r;
        return 0;
    }
    case GO_BOX: { /* cpu_evaluators.go:28-36; fields dims(3), round */
        float dx = 0.5f * f[0], dy = 0.5f * f[1], dz = 0.5f * f[2], r = f[3];
        for (size_t i = 0; i < n; i++) {
            float qx = (go_abs(pos[i].x) - dx) + r, qy = (go_abs(pos[i].y) - dy) + r, qz = (go_abs(pos[i].z) - dz) + r;
            dist[i] = norm3(go_max(qx, 0), go_max(qy, 0), go_max(qz, 0)) + go_min(go_max(qx, go_max(qy, qz)), 0.0f) - r;
        }
        return 0;
    }
    case GO_BOXFRAME: { /* cpu_evaluators.go:38-57, args primitives.go:292-297; fields dims(3), e */
        float e = f[3];
        float bx = 0.5f * f[0] + (-2 * e), by = 0.5f * f[1] + (-2 * e), bz = 0.5f * f[2] + (-2 * e);
        for (size_t i = 0; i < n; i++) {
            float px = go_abs(pos[i].x) - bx, py = go_abs(pos[i].y) - by, pz = go_abs(pos[i].z) - bz;
            float qx = go_abs(px + e) + (-e), qy = go_abs(py + e) + (-e), qz = go_abs(pz + e) + (-e);
            float s1 = go_min(0, go_max(px, go_max(qy, qz)));
            float n1 = norm3(go_max(px, 0), go_max(qy, 0), go_max(qz, 0)) + s1;
            float s2 = go_min(0, go_max(qx, go_max(py, qz)));
            float n2 = norm3(go_max(qx, 0), go_max(py, 0), go_max(qz, 0)) + s2;
            float s3 = go_min(0, go_max(qx, go_max(qy, pz)));
            float n3 = norm3(go_max(qx, 0), go_max(qy, 0), go_max(pz, 0)) + s3;
            dist[i] = go_min(n1, go_min(n2, n3));
        }
        return 0;
    }
    case GO_TORUS: { /* cpu_evaluators.go:59-68; fields rLesser, rGreater */
        float t2 = f[0], t1 = f[1];
        for (size_t i = 0; i < n; i++) {
            /* p = (x, z, y); q = (hypot(p.x, p.z) - t1, p.y) */
            float qx = go_hypot(pos[i].x, pos[i].y) - t1, qy = pos[i].z;
            dist[i] = norm2(qx, qy) - t2;
        }
        return 0;
    }
    case GO_CYLINDER: { /* cpu_evaluators.go:70-88, args primitives.go:147-149; fields r, h, round */
        float r = f[0], round = f[2], h = (f[1] - 2 * round) / 2;
        if (round == 0) {
            for (size_t i = 0; i < n; i++) {
                float dx = go_hypot(pos[i].x, pos[i].y) - r;
                float dy = go_abs(pos[i].z) - h;
                dist[i] = go_min(0, go_max(dx, dy)) + go_hypot(go_max(0, dx), go_max(0, dy));
            }
        } else {
            for (size_t i = 0; i < n; i++) {
                float dx = go_hypot(pos[i].x, pos[i].y) - r + round;
                float dy = go_abs(pos[i].z) - h;
                dist[i] = go_min(go_max(dx, dy), 0) + go_hypot(go_max(dx, 0), go_max(dy, 0)) - round;
            }
        }
        return 0;
    }
    case GO_HEX: { /* cpu_evaluators.go:90-105; fields side, h */
        const float k1 = (float)(-GSDF_TRIBISECT), k2 = 0.5f, k3 = 0.57735f;
        const float two_k1 = (float)(2 * -GSDF_TRIBISECT), two_k2 = 1.0f; /* Go folds 2*k1 as a constant */
        float h1 = f[0], h2 = f[1];
        float clm = k3 * h1;
        for (size_t i = 0; i < n; i++) {
            float px = go_abs(pos[i].x), py = go_abs(pos[i].y), pz = go_abs(pos[i].z);
            float pm = go_min(k1 * px + k2 * py, 0);
            px -= two_k1 * pm;
            py -= two_k2 * pm;
            float d1 = go_hypot(px - clampf(px, -clm, clm), py - h1) * signf(py - h1);
            float d2 = pz - h2;
            dist[i] = go_min(go_max(d1, d2), 0) + go_hypot(go_max(d1, 0), go_max(d2, 0));
        }
        return 0;
    }
    case GO_UNION: { /* cpu_evaluators.go:124-144 */
        if (nd->nchild < 2) return -1;
        SCRATCH(float, aux, n);
        err = eval3(t, child_id(t, nd, 0), pos, dist, n);
        for (int c = 1; c < nd->nchild && !err; c++) {
            err = eval3(t, child_id(t, nd, c), pos, aux, n);
            if (!err) for (size_t i = 0; i < n; i++) dist[i] = go_min(dist[i], aux[i]);
        }
        free(aux);
        return err;
    }
    case GO_INTERSECT: case GO_DIFF: case GO_XOR:
    case GO_SMOOTH_UNION: case GO_SMOOTH_DIFF: case GO_SMOOTH_INTERSECT: { /* cpu_evaluators.go:146-286 */
        if (nd->nchild != 2) return -1;
        SCRATCH(float, d2, n);
        err = eval3(t, child_id(t, nd, 0), pos, dist, n);
        if (!err) err = eval3(t, child_id(t, nd, 1), pos, d2, n);
        if (!err) {
            float k = f[0];
            for (size_t i = 0; i < n; i++) {
                float a = dist[i], b = d2[i], h;
                switch (nd->kind) {
                case GO_INTERSECT: dist[i] = go_max(a, b); break;
                case GO_DIFF: dist[i] = go_max(a, -b); break;
                case GO_XOR: dist[i] = go_max(go_min(a, b), -go_max(a, b)); break;
                case GO_SMOOTH_UNION:
                    h = clampf(0.5f + 0.5f * (b - a) / k, 0, 1);
                    dist[i] = mixf(b, a, h) - k * h * (1 - h);
                    break;
                case GO_SMOOTH_DIFF:
                    h = clampf(0.5f - 0.5f * (b + a) / k, 0, 1);
                    dist[i] = mixf(a, -b, h) + k * h * (1 - h);
                    break;
                default: /* GO_SMOOTH_INTERSECT */
                    h = clampf(0.5f - 0.5f * (b - a) / k, 0, 1);
                    dist[i] = mixf(b, a, h) + k * h * (1 - h);
                    break;
                }
            }
        }
        free(d2);
        return err;
    }
    case GO_SCALE: { /* cpu_evaluators.go:288-312; field scale */
        SCRATCH(v3, sc, n);
        float factor = f[0], inv = 1.f / f[0];
        for (size_t i = 0; i < n; i++) { sc[i].x = inv * pos[i].x; sc[i].y = inv * pos[i].y; sc[i].z = inv * pos[i].z; }
        err = eval3(t, child_id(t, nd, 0), sc, dist, n);
        if (!err) for (size_t i = 0; i < n; i++) dist[i] *= factor;
        free(sc);
        return err;
    }
    case GO_SYMMETRY: { /* cpu_evaluators.go:314-343; iparam[0] = xyz bits */
        SCRATCH(v3, tr, n);
        int xb = nd->iparam[0] & 1, yb = nd->iparam[0] & 2, zb = nd->iparam[0] & 4;
        for (size_t i = 0; i < n; i++) {
            tr[i] = pos[i];
            if (xb) tr[i].x = go_abs(pos[i].x);
            if (yb) tr[i].y = go_abs(pos[i].y);
            if (zb) tr[i].z = go_abs(pos[i].z);
        }
        err = eval3(t, child_id(t, nd, 0), tr, dist, n);
        free(tr);
        return err;
    }
    case GO_ARRAY: { /* cpu_evaluators.go:345-397; fields d(3); iparam nx,ny,nz */
        SCRATCH(v3, tr, n);
        float *aux = (float *)malloc(sizeof(float) * (n ? n : 1));
        if (!aux) { free(tr); return -2; }
        float sx = f[0], sy = f[1], sz = f[2];
        float nx = (float)nd->iparam[0] + -1, ny = (float)nd->iparam[1] + -1, nz = (float)nd->iparam[2] + -1;
        for (size_t i = 0; i < n; i++) dist[i] = GSDF_LARGENUM;
        for (int k = 0; k < 2 && !err; k++)
            for (int j = 0; j < 2 && !err; j++)
                for (int ii = 0; ii < 2 && !err; ii++) {
                    float fi = (float)ii, fj = (float)j, fk = (float)k;
                    for (size_t ip = 0; ip < n; ip++) {
                        v3 p = pos[ip];
                        float idx = go_round(p.x / sx), idy = go_round(p.y / sy), idz = go_round(p.z / sz);
                        float ox = signf(p.x - sx * idx), oy = signf(p.y - sy * idy), oz = signf(p.z - sz * idz);
                        float rx = clampf(idx + fi * ox, 0, nx), ry = clampf(idy + fj * oy, 0, ny), rz = clampf(idz + fk * oz, 0, nz);
                        tr[ip].x = p.x - sx * rx; tr[ip].y = p.y - sy * ry; tr[ip].z = p.z - sz * rz;
                    }
                    err = eval3(t, child_id(t, nd, 0), tr, aux, n);
                    if (!err) for (size_t ip = 0; ip < n; ip++) dist[ip] = go_min(dist[ip], aux[ip]);
                }
        free(aux); free(tr);
        return err;
    }
    case GO_ELONGATE: { /* cpu_evaluators.go:399-426; fields h(3) */
        SCRATCH(v3, tr, n);
        float *aux = (float *)malloc(sizeof(float) * (n ? n : 1));
        if (!aux) { free(tr); return -2; }
        float hx = 0.5f * f[0], hy = 0.5f * f[1], hz = 0.5f * f[2];
        for (size_t i = 0; i < n; i++) {
            float qx = go_abs(pos[i].x) - hx, qy = go_abs(pos[i].y) - hy, qz = go_abs(pos[i].z) - hz;
            aux[i] = go_min(go_max(qx, go_max(qy, qz)), 0);
            tr[i].x = go_max(qx, 0); tr[i].y = go_max(qy, 0); tr[i].z = go_max(qz, 0);
        }
        err = eval3(t, child_id(t, nd, 0), tr, dist, n);
        if (!err) for (size_t i = 0; i < n; i++) dist[i] += aux[i];
        free(aux); free(tr);
        return err;
    }
    case GO_SHELL: { /* cpu_evaluators.go:428-452; field thick */
        SCRATCH(v3, tr, n);
        float th = f[0], inv = 1 / th;
        for (size_t i = 0; i < n; i++) { tr[i].x = inv * pos[i].x; tr[i].y = inv * pos[i].y; tr[i].z = inv * pos[i].z; }
        err = eval3(t, child_id(t, nd, 0), tr, dist, n);
        if (!err) for (size_t i = 0; i < n; i++) dist[i] = th * (go_abs(dist[i]) - th);
        free(tr);
        return err;
    }
    case GO_OFFSET: { /* cpu_evaluators.go:454-468; field off */
        err = eval3(t, child_id(t, nd, 0), pos, dist, n);
        if (!err) for (size_t i = 0; i < n; i++) dist[i] = dist[i] + f[0];
        return err;
    }
    case GO_TRANSLATE: { /* cpu_evaluators.go:470-486; fields p(3) */
        SCRATCH(v3, tr, n);
        for (size_t i = 0; i < n; i++) { tr[i].x = pos[i].x - f[0]; tr[i].y = pos[i].y - f[1]; tr[i].z = pos[i].z - f[2]; }
        err = eval3(t, child_id(t, nd, 0), tr, dist, n);
        free(tr);
        return err;
    }
    case GO_TRANSFORM: { /* cpu_evaluators.go:488-504; fields tInv rows 0..2 (12 floats, row-major, w row dropped).
                          * ms3.Mat4.MulPosition (UNPINNED): x' = x00*x + x01*y + x02*z + x03, etc. */
        SCRATCH(v3, tr, n);
        for (size_t i = 0; i < n; i++) {
            v3 p = pos[i];
            tr[i].x = f[0] * p.x + f[1] * p.y + f[2] * p.z + f[3];
            tr[i].y = f[4] * p.x + f[5] * p.y + f[6] * p.z + f[7];
            tr[i].z = f[8] * p.x + f[9] * p.y + f[10] * p.z + f[11];
        }
        err = eval3(t, child_id(t, nd, 0), tr, dist, n);
        free(tr);
        return err;
    }
    case GO_CIRCARRAY: { /* cpu_evaluators.go:1042-1092; iparam nInst, circleDiv */
        SCRATCH(v3, p0, n);
        v3 *p1 = (v3 *)malloc(sizeof(v3) * (n ? n : 1));
        float *d1 = (float *)malloc(sizeof(float) * (n ? n : 1));
        if (!p1 || !d1) { free(p0); free(p1); free(d1); return -2; }
        float ncirc = (float)nd->iparam[1];
        float angle = (float)(2 * GO_PI) / ncirc;
        float ninsm1 = (float)(nd->iparam[0] - 1);
        for (size_t i = 0; i < n; i++) {
            v3 p = pos[i];
            float pangle = go_atan2(p.y, p.x);
            float idf = go_floor(pangle / angle);
            if (idf < 0) idf += ncirc;
            float i0, i1;
            if (idf >= ninsm1) { i0 = ninsm1; i1 = 0; } else { i0 = idf; i1 = idf + 1; }
            /* ms2.MulMatVecTrans(RotationMat2(a), p) = (c*x + s*y, -s*x + c*y)  (UNPINNED helper) */
            float s0 = go_sin(angle * i0), c0 = go_cos(angle * i0);
            float s1 = go_sin(angle * i1), c1 = go_cos(angle * i1);
            p0[i].x = c0 * p.x + s0 * p.y; p0[i].y = -s0 * p.x + c0 * p.y; p0[i].z = p.z;
            p1[i].x = c1 * p.x + s1 * p.y; p1[i].y = -s1 * p.x + c1 * p.y; p1[i].z = p.z;
        }
        err = eval3(t, child_id(t, nd, 0), p1, d1, n);
        if (!err) err = eval3(t, child_id(t, nd, 0), p0, dist, n);
        if (!err) for (size_t i = 0; i < n; i++) dist[i] = go_min(dist[i], d1[i]);
        free(p0); free(p1); free(d1);
        return err;
    }
    case GO_TWIST: { /* cpu_evaluators.go:1257-1274; field k */
        SCRATCH(v3, tr, n);
        for (size_t i = 0; i < n; i++) {
            v3 p = pos[i];
            float c = go_cos(f[0] * p.z), s = go_sin(f[0] * p.z);
            tr[i].x = c * p.x - s * p.y; tr[i].y = s * p.x + c * p.y; tr[i].z = p.z;
        }
        err = eval3(t, child_id(t, nd, 0), tr, dist, n);
        free(tr);
        return err;
    }
    case GO_BOUNDS3: /* glbuild/glbuild.go:1095-1101: overloadBounds3.Evaluate forwards to the wrapped shader */
        return eval3(t, child_id(t, nd, 0), pos, dist, n);
    case GO_EXTRUDE: { /* cpu_evaluators.go:506-531; field h */
        SCRATCH(v2, p2, n);
        for (size_t i = 0; i < n; i++) { p2[i].x = pos[i].x; p2[i].y = pos[i].y; }
        err = eval2(t, child_id(t, nd, 0), p2, dist, n);
        if (!err) {
            float h = f[0] / 2;
            for (size_t i = 0; i < n; i++) {
                float d = dist[i], wy = go_abs(pos[i].z) - h;
                dist[i] = go_min(0, go_max(d, wy)) + go_hypot(go_max(d, 0), go_max(wy, 0));
            }
        }
        free(p2);
        return err;
    }
    case GO_REVOLVE: { /* cpu_evaluators.go:533-549; field off */
        SCRATCH(v2, p2, n);
        for (size_t i = 0; i < n; i++) { p2[i].x = go_hypot(pos[i].x, pos[i].z) - f[0]; p2[i].y = pos[i].y; }
        err = eval2(t, child_id(t, nd, 0), p2, dist, n);
        free(p2);
        return err;
    }
    case GO_SCREW: { /* forge/threads/threads.go:141-181,198-202; fields pitch, lead, lengthDiv2, taper */
        SCRATCH(v2, tr, n);
        float pitch = f[0], lead = f[1], L = f[2], taper = f[3];
        float tanTaper = go_tan(taper);
        for (size_t i = 0; i < n; i++) {
            v3 p = pos[i];
            float y = go_hypot(p.x, p.y);
            y += p.z * tanTaper;
            float theta = go_atan2(p.y, p.x);
            float z = p.z + lead * theta / (float)(2 * GO_PI);
            /* sawTooth(z, pitch) */
            float x = z + pitch / 2;
            float tt = x / pitch;
            tr[i].x = pitch * (tt - go_floor(tt)) - pitch / 2;
            tr[i].y = y;
        }
        err = eval2(t, child_id(t, nd, 0), tr, dist, n);
        if (!err) for (size_t i = 0; i < n; i++) dist[i] = go_max(dist[i], go_abs(pos[i].z) - L);
        free(tr);
        return err;
    }
    default:
        return -1;
    }
}

static int eval2(const go_tree *t, int id, const v2 *pos, float *dist, size_t n) {
    const go_node *nd = node_at(t, id);
    if (!nd) return -1;
    const float *f = nd->fparam;
    const float *aux = t->aux + nd->aux_off;
    int err = 0;
    switch (nd->kind) {
    case GO_BOUNDS2: /* glbuild/glbuild.go:1120-1126 */
        return eval2(t, child_id(t, nd, 0), pos, dist, n);
    case GO_CIRCLE2D: /* cpu_evaluators.go:661-667 */
        for (size_t i = 0; i < n; i++) dist[i] = norm2(pos[i].x, pos[i].y) - f[0];
        return 0;
    case GO_RECT2D: { /* cpu_evaluators.go:685-692; fields d(2) */
        float bx = 0.5f * f[0], by = 0.5f * f[1];
        for (size_t i = 0; i < n; i++) {
            float dx = go_abs(pos[i].x) - bx, dy = go_abs(pos[i].y) - by;
            dist[i] = norm2(go_max(dx, 0), go_max(dy, 0)) + go_min(0, go_max(dx, dy));
        }
        return 0;
    }
    case GO_LINE2D: { /* cpu_evaluators.go:551-562; fields width, a(2), b(2) */
        float ax = f[1], ay = f[2], bax = f[3] - f[1], bay = f[4] - f[2];
        float dotba = bax * bax + bay * bay, w = f[0] / 2;
        for (size_t i = 0; i < n; i++) {
            float pax = pos[i].x - ax, pay = pos[i].y - ay;
            float h = clampf((pax * bax + pay * bay) / dotba, 0, 1);
            dist[i] = norm2(pax - h * bax, pay - h * bay) - w;
        }
        return 0;
    }
    case GO_LINES2D: { /* cpu_evaluators.go:1145-1160; field width; aux = segments (ax,ay,bx,by)* */
        float w = f[0] / 2;
        int nseg = nd->aux_cnt / 4;
        for (size_t i = 0; i < n; i++) {
            float d = 1e23f;
            for (int s = 0; s < nseg; s++) {
                float ax = aux[4 * s], ay = aux[4 * s + 1], bx = aux[4 * s + 2], by = aux[4 * s + 3];
                float pax = pos[i].x - ax, pay = pos[i].y - ay, bax = bx - ax, bay = by - ay;
                float dotba = bax * bax + bay * bay;
                float h = clampf((pax * bax + pay * bay) / dotba, 0, 1);
                float ex = pax - h * bax, ey = pay - h * bay;
                d = go_min(d, ex * ex + ey * ey);
            }
            dist[i] = sqrtf(d) - w;
        }
        return 0;
    }
    case GO_ARC2D: { /* cpu_evaluators.go:564-579; fields radius, angle, thick */
        float r = f[0], th = f[2] / 2;
        float s = go_sin(f[1] / 2), c = go_cos(f[1] / 2);
        float scrx = r * s, scry = r * c;
        for (size_t i = 0; i < n; i++) {
            float px = go_abs(pos[i].x), py = pos[i].y;
            if (c * px > s * py) dist[i] = norm2(px - scrx, py - scry) - th;
            else dist[i] = go_abs(norm2(px, py) - r) - th;
        }
        return 0;
    }
    case GO_EQTRI2D: { /* cpu_evaluators.go:669-683; field hTri */
        const float k = (float)GSDF_SQRT3;
        float r = f[0] / k;
        for (size_t i = 0; i < n; i++) {
            float px = go_abs(pos[i].x) - r, py = pos[i].y + r / k;
            if (px + k * py > 0) {
                float nx = px - k * py, ny = -k * px - py;
                px = 0.5f * nx; py = 0.5f * ny;
            }
            px -= clampf(px, -2 * r, 0);
            dist[i] = -norm2(px, py) * signf(py);
        }
        return 0;
    }
    case GO_HEX2D: { /* cpu_evaluators.go:718-729; field side */
        const float kx = (float)(-GSDF_TRIBISECT), ky = 0.5f, kz = 0.577350269f;
        float r = f[0];
        for (size_t i = 0; i < n; i++) {
            float px = go_abs(pos[i].x), py = go_abs(pos[i].y);
            float m = 2 * go_min(kx * px + ky * py, 0);
            px = px - m * kx; py = py - m * ky;
            px = px - clampf(px, -kz * r, kz * r); py = py - r;
            dist[i] = signf(py) * norm2(px, py);
        }
        return 0;
    }
    case GO_OCT2D: { /* cpu_evaluators.go:731-748; field c */
        const float kx = -0.9238795325f, ky = 0.3826834323f, kz = 0.4142135623f;
        float r = f[0], kzr = kz * r, nkzr = -kzr;
        for (size_t i = 0; i < n; i++) {
            float px = go_abs(pos[i].x), py = go_abs(pos[i].y);
            float m = 2 * go_min(kx * px + ky * py, 0);
            px = px - m * kx; py = py - m * ky;
            m = 2 * go_min(-kx * px + ky * py, 0);
            px = px - m * -kx; py = py - m * ky;
            px = px - clampf(px, nkzr, kzr); py = py - r;
            dist[i] = signf(py) * norm2(px, py);
        }
        return 0;
    }
    case GO_DIAMOND2D: { /* cpu_evaluators.go:694-703; fields d(2) */
        float bx = 0.5f * f[0], by = 0.5f * f[1];
        for (size_t i = 0; i < n; i++) {
            float px = go_abs(pos[i].x), py = go_abs(pos[i].y);
            float ux = bx - 2 * px, uy = by - 2 * py;
            float h = clampf((ux * bx - uy * by) / (bx * bx + by * by), -1, 1);
            float d = norm2(px - (0.5f * bx) * (1 - h), py - (0.5f * by) * (1 + h));
            dist[i] = d * signf(px * by + py * bx - bx * by);
        }
        return 0;
    }
    case GO_ROUNDX2D: { /* cpu_evaluators.go:705-716; fields dim, thick */
        for (size_t i = 0; i < n; i++) {
            float px = go_abs(pos[i].x), py = go_abs(pos[i].y);
            float sub = 0.5f * go_min(px + py, f[0]);
            dist[i] = norm2(px - sub, py - sub) - f[1];
        }
        return 0;
    }
    case GO_POLY2D: { /* cpu_evaluators.go:793-818; aux = vertices (x,y)* */
        int nv = nd->aux_cnt / 2;
        if (nv < 3) return -1;
        const v2 *verts = (const v2 *)aux;
        for (size_t i = 0; i < n; i++) {
            v2 p = pos[i];
            float d0x = p.x - verts[0].x, d0y = p.y - verts[0].y;
            float d = d0x * d0x + d0y * d0y;
            float s = 1.0f;
            int jv = nv - 1;
            for (int iv = 0; iv < nv; iv++) {
                v2 v1 = verts[iv], v2_ = verts[jv];
                float ex = v2_.x - v1.x, ey = v2_.y - v1.y;
                float wx = p.x - v1.x, wy = p.y - v1.y;
                float c = clampf((wx * ex + wy * ey) / (ex * ex + ey * ey), 0, 1);
                float bx = wx - c * ex, by = wy - c * ey;
                d = go_min(d, bx * bx + by * by);
                int b1 = p.y >= v1.y, b2 = p.y < v2_.y, b3 = ex * wy > ey * wx;
                if ((b1 && b2 && b3) || (!b1 && !b2 && !b3)) s = -s;
                jv = iv;
            }
            dist[i] = s * sqrtf(d);
        }
        return 0;
    }
    case GO_UNION2D: { /* cpu_evaluators.go:821-845 */
        if (nd->nchild < 2) return -1;
        SCRATCH(float, a, n);
        err = eval2(t, child_id(t, nd, 0), pos, dist, n);
        for (int c = 1; c < nd->nchild && !err; c++) {
            err = eval2(t, child_id(t, nd, c), pos, a, n);
            if (!err) for (size_t i = 0; i < n; i++) dist[i] = go_min(dist[i], a[i]);
        }
        free(a);
        return err;
    }
    case GO_INTERSECT2D: case GO_DIFF2D: case GO_XOR2D: { /* cpu_evaluators.go:847-912 */
        if (nd->nchild != 2) return -1;
        SCRATCH(float, d2, n);
        err = eval2(t, child_id(t, nd, 0), pos, dist, n);
        if (!err) err = eval2(t, child_id(t, nd, 1), pos, d2, n);
        if (!err) for (size_t i = 0; i < n; i++) {
            float a = dist[i], b = d2[i];
            dist[i] = nd->kind == GO_INTERSECT2D ? go_max(a, b)
                    : nd->kind == GO_DIFF2D      ? go_max(a, -b)
                                                 : go_max(go_min(a, b), -go_max(a, b));
        }
        free(d2);
        return err;
    }
    case GO_ARRAY2D: { /* cpu_evaluators.go:914-962; fields d(2); iparam nx, ny */
        SCRATCH(v2, tr, n);
        float *a = (float *)malloc(sizeof(float) * (n ? n : 1));
        if (!a) { free(tr); return -2; }
        float sx = f[0], sy = f[1];
        float nx = (float)nd->iparam[0] + -1, ny = (float)nd->iparam[1] + -1;
        for (size_t i = 0; i < n; i++) dist[i] = GSDF_LARGENUM;
        for (int j = 0; j < 2 && !err; j++)
            for (int ii = 0; ii < 2 && !err; ii++) {
                float fi = (float)ii, fj = (float)j;
                for (size_t ip = 0; ip < n; ip++) {
                    v2 p = pos[ip];
                    float idx = go_round(p.x / sx), idy = go_round(p.y / sy);
                    float ox = signf(p.x - sx * idx), oy = signf(p.y - sy * idy);
                    float rx = clampf(idx + fi * ox, 0, nx), ry = clampf(idy + fj * oy, 0, ny);
                    tr[ip].x = p.x - sx * rx; tr[ip].y = p.y - sy * ry;
                }
                err = eval2(t, child_id(t, nd, 0), tr, a, n);
                if (!err) for (size_t ip = 0; ip < n; ip++) dist[ip] = go_min(dist[ip], a[ip]);
            }
        free(a); free(tr);
        return err;
    }
    case GO_OFFSET2D: /* cpu_evaluators.go:964-978; field f */
        err = eval2(t, child_id(t, nd, 0), pos, dist, n);
        if (!err) for (size_t i = 0; i < n; i++) dist[i] = dist[i] + f[0];
        return err;
    case GO_TRANSLATE2D: { /* cpu_evaluators.go:980-996; fields p(2) */
        SCRATCH(v2, tr, n);
        for (size_t i = 0; i < n; i++) { tr[i].x = pos[i].x - f[0]; tr[i].y = pos[i].y - f[1]; }
        err = eval2(t, child_id(t, nd, 0), tr, dist, n);
        free(tr);
        return err;
    }
    case GO_SYMMETRY2D: { /* cpu_evaluators.go:998-1024 */
        SCRATCH(v2, tr, n);
        int xb = nd->iparam[0] & 1, yb = nd->iparam[0] & 2;
        for (size_t i = 0; i < n; i++) {
            tr[i] = pos[i];
            if (xb) tr[i].x = go_abs(pos[i].x);
            if (yb) tr[i].y = go_abs(pos[i].y);
        }
        err = eval2(t, child_id(t, nd, 0), tr, dist, n);
        free(tr);
        return err;
    }
    case GO_ANNULUS2D: /* cpu_evaluators.go:1026-1040; field r */
        err = eval2(t, child_id(t, nd, 0), pos, dist, n);
        if (!err) for (size_t i = 0; i < n; i++) dist[i] = go_abs(dist[i]) - f[0];
        return err;
    case GO_CIRCARRAY2D: { /* cpu_evaluators.go:1094-1143 */
        SCRATCH(v2, p0, n);
        v2 *p1 = (v2 *)malloc(sizeof(v2) * (n ? n : 1));
        float *d1 = (float *)malloc(sizeof(float) * (n ? n : 1));
        if (!p1 || !d1) { free(p0); free(p1); free(d1); return -2; }
        float ncirc = (float)nd->iparam[1];
        float angle = (float)(2 * GO_PI) / ncirc;
        float ninsm1 = (float)(nd->iparam[0] - 1);
        for (size_t i = 0; i < n; i++) {
            v2 p = pos[i];
            float idf = go_floor(go_atan2(p.y, p.x) / angle);
            if (idf < 0) idf += ncirc;
            float i0, i1;
            if (idf >= ninsm1) { i0 = ninsm1; i1 = 0; } else { i0 = idf; i1 = idf + 1; }
            float s0 = go_sin(angle * i0), c0 = go_cos(angle * i0);
            float s1 = go_sin(angle * i1), c1 = go_cos(angle * i1);
            p0[i].x = c0 * p.x + s0 * p.y; p0[i].y = -s0 * p.x + c0 * p.y;
            p1[i].x = c1 * p.x + s1 * p.y; p1[i].y = -s1 * p.x + c1 * p.y;
        }
        err = eval2(t, child_id(t, nd, 0), p1, d1, n);
        if (!err) err = eval2(t, child_id(t, nd, 0), p0, dist, n);
        if (!err) for (size_t i = 0; i < n; i++) dist[i] = go_min(dist[i], d1[i]);
        free(p0); free(p1); free(d1);
        return err;
    }
    case GO_TRANSLATEMULTI2D: { /* cpu_evaluators.go:1162-1184; aux = displacements (x,y)* */
        SCRATCH(v2, tr, n);
        float *d1 = (float *)malloc(sizeof(float) * (n ? n : 1));
        if (!d1) { free(tr); return -2; }
        int nd_ = nd->aux_cnt / 2;
        for (size_t i = 0; i < n; i++) dist[i] = 3.40282346638528859811704183484516925440e+38f;
        for (int k = 0; k < nd_ && !err; k++) {
            for (size_t i = 0; i < n; i++) { tr[i].x = pos[i].x - aux[2 * k]; tr[i].y = pos[i].y - aux[2 * k + 1]; }
            err = eval2(t, child_id(t, nd, 0), tr, d1, n);
            if (!err) for (size_t i = 0; i < n; i++) dist[i] = go_min(dist[i], d1[i]);
        }
        free(d1); free(tr);
        return err;
    }
    case GO_ROTATE2D: { /* cpu_evaluators.go:1186-1203; fields tInv (x00,x01,x10,x11). ms2.MulMatVec (UNPINNED) */
        SCRATCH(v2, tr, n);
        for (size_t i = 0; i < n; i++) {
            tr[i].x = f[0] * pos[i].x + f[1] * pos[i].y;
            tr[i].y = f[2] * pos[i].x + f[3] * pos[i].y;
        }
        err = eval2(t, child_id(t, nd, 0), tr, dist, n);
        free(tr);
        return err;
    }
    case GO_SCALE2D: { /* cpu_evaluators.go:1205-1226; field scale */
        SCRATCH(v2, tr, n);
        float inv = 1.f / f[0];
        for (size_t i = 0; i < n; i++) { tr[i].x = inv * pos[i].x; tr[i].y = inv * pos[i].y; }
        err = eval2(t, child_id(t, nd, 0), tr, dist, n);
        if (!err) for (size_t i = 0; i < n; i++) dist[i] = dist[i] * f[0];
        free(tr);
        return err;
    }
    case GO_ELONGATE2D: { /* cpu_evaluators.go:1228-1255; fields h(2) */
        SCRATCH(v2, tr, n);
        float *a = (float *)malloc(sizeof(float) * (n ? n : 1));
        if (!a) { free(tr); return -2; }
        float hx = 0.5f * f[0], hy = 0.5f * f[1];
        for (size_t i = 0; i < n; i++) {
            float qx = go_abs(pos[i].x) - hx, qy = go_abs(pos[i].y) - hy;
            a[i] = go_min(go_max(qx, qy), 0);
            tr[i].x = go_max(qx, 0); tr[i].y = go_max(qy, 0);
        }
        err = eval2(t, child_id(t, nd, 0), tr, dist, n);
        if (!err) for (size_t i = 0; i < n; i++) dist[i] += a[i];
        free(a); free(tr);
        return err;
    }
    case GO_ELLIPSE2D: { /* cpu_evaluators.go:750-791; fields a, b */
        const float sq3 = (float)GSDF_SQRT3;
        for (size_t i = 0; i < n; i++) {
            float a = f[0], b = f[1];
            float px = go_abs(pos[i].x), py = go_abs(pos[i].y);
            if (px > py) { float t = px; px = py; py = t; t = a; a = b; b = t; }
            float l = b * b - a * a;
            float m = a * px / l, m2 = m * m;
            float nn = b * py / l, n2 = nn * nn;
            float c = (m2 + n2 - 1) / 3;
            float c3 = c * c * c;
            float q = c3 + 2 * m2 * n2;
            float d = c3 + m2 * n2;
            float g = m + m * n2;
            float co;
            if (d < 0) {
                float h = go_acos(q / c3) / 3;
                float sh = go_sin(h), ch = go_cos(h);
                float t = sq3 * sh;
                float rx = sqrtf(-c * (ch + t + 2) + m2);
                float ry = sqrtf(-c * (ch - t + 2) + m2);
                co = (ry + signf(l) * rx + go_abs(g) / (rx * ry) - m) / 2;
            } else {
                float h = 2 * m * nn * sqrtf(d);
                float s_ = signf(q + h) * go_cbrt(go_abs(q + h));
                float u = signf(q - h) * go_cbrt(go_abs(q - h));
                float rx = -s_ - u - 4 * c + 2 * m2;
                float ry = sq3 * (s_ - u);
                float rm = go_hypot(rx, ry);
                co = (ry / sqrtf(rm - rx) + 2 * g / rm - m) / 2;
            }
            float rx2 = a * co, ry2 = b * sqrtf(1 - co * co);
            dist[i] = norm2(rx2 - px, ry2 - py) * signf(py - ry2);
        }
        return 0;
    }
    case GO_BEZIERQ2D: { /* cpu_evaluators.go:581-659; fields a(2), b(2), c(2), thick */
        const float sq3 = (float)GSDF_SQRT3;
        float thick = f[6] / 2;
        float Ax = f[0], Ay = f[1], Bx = f[2], By = f[3], Cx = f[4], Cy = f[5];
        float ax = Bx - Ax, ay = By - Ay;
        float a2 = ax * ax + ay * ay;
        float bx = Ax + (Cx - 2 * Bx), by = Ay + (Cy - 2 * By);
        float cx = 2 * ax, cy = 2 * ay;
        float kk = 1.f / (bx * bx + by * by);
        float kx = kk * (ax * bx + ay * by);
        float kx2 = kx * kx;
        for (size_t i = 0; i < n; i++) {
            float dx = Ax - pos[i].x, dy = Ay - pos[i].y;
            float ky = kk * (2 * a2 + (dx * bx + dy * by)) / 3;
            float kz = kk * (dx * ax + dy * ay);
            float g = ky - kx2;
            float q = kx * (2 * kx2 - 3 * ky) + kz;
            float g3 = g * g * g;
            float q2 = q * q;
            float h = q2 + 4 * g3;
            float res;
            if (h >= 0) {
                h = sqrtf(h);
                float xx = 0.5f * (h + -q), xy = 0.5f * (-h + -q);
                if (go_abs(g) < 0.001f) {
                    float k = (1.0f - g3 / q2) * g3 / q;
                    xx = k; xy = -k - q;
                }
                float ux = signf(xx) * go_pow_frac(go_abs(xx), (float)(1. / 3));
                float uy = signf(xy) * go_pow_frac(go_abs(xy), (float)(1. / 3));
                float t = ux + uy;
                t -= (t * (t * t + 3.0f * g) + q) / (3.0f * t * t + 3.0f * g);
                t = clampf(t - kx, 0, 1);
                float wx = dx + t * (cx + t * bx), wy = dy + t * (cy + t * by);
                res = wx * wx + wy * wy;
            } else {
                float z = sqrtf(-g);
                float xm = sqrtf(0.5f + 0.5f * (q / (2 * g * z))); /* cos_acos_3, gsdf.go:186-189 */
                float m = xm * (xm * (xm * (xm * -0.008972f + 0.039071f) - 0.107074f) + 0.576975f) + 0.5f;
                float nn = sqrtf(1 - m * m);
                nn *= sq3;
                float tx = clampf((m + m) * z - kx, 0, 1);
                float ty = clampf((-nn - m) * z - kx, 0, 1);
                float qxx = dx + tx * (cx + tx * bx), qxy = dy + tx * (cy + tx * by);
                float qyx = dx + ty * (cx + ty * bx), qyy = dy + ty * (cy + ty * by);
                float ddx = qxx * qxx + qxy * qxy, ddy = qyx * qyx + qyy * qyy;
                res = ddx < ddy ? ddx : ddy;
            }
            dist[i] = sqrtf(res) - thick;
        }
        return 0;
    }
    default:
        return -1;
    }
}

int go_eval3(const go_tree *t, const float *pos_xyz, float *dist, size_t n) {
    return eval3(t, t->root, (const v3 *)pos_xyz, dist, n);
}
int go_eval2(const go_tree *t, const float *pos_xy, float *dist, size_t n) {
    return eval2(t, t->root, (const v2 *)pos_xy, dist, n);
}

/* ------------------------------------------------------------------------------------------------------------
 * FlatRenderer (glrender/flatrenderer.go)
 * ---------------------------------------------------------------------------------------------------------- */
static void scale_centered_101(const float bbmin[3], const float bbmax[3], float omin[3], float omax[3]) {
    /* ms3.Box.ScaleCentered(1.01) (UNPINNED helper): size*=s about the centre.
     * Restated as: c = (min+max)*0.5 ; half = (max-min)*s*0.5 ; [c-half, c+half]. Reproduces README's 281x281x85. */
    for (int a = 0; a < 3; a++) {
        float size = bbmax[a] - bbmin[a];
        float ns = 1.01f * size;
        float c = bbmin[a] + size * 0.5f; /* Box.Center(): Min + Size/2 */
        float half = ns * 0.5f;
        omin[a] = c - half;
        omax[a] = c + half;
    }
}

int go_flat_lattice(const float bbmin[3], const float bbmax[3], float res, go_lattice *out) {
    if (!(res > 0)) return -1;
    float mn[3], mx[3];
    scale_centered_101(bbmin, bbmax, mn, mx);
    for (int a = 0; a < 3; a++) {
        float sz = mx[a] - mn[a];
        int n = (int)ceilf(sz / res); /* flatrenderer.go:50-52 */
        if (n <= 0) return -1;
        out->n[a] = n;
        out->origin[a] = mn[a];
    }
    out->res = res;
    return 0;
}

int go_octree_levels(const float bbmin[3], const float bbmax[3], float res) {
    if (!(res > 0) || isinf(res)) return -1;
    float mn[3], mx[3];
    scale_centered_101(bbmin, bbmax, mn, mx);
    float longAxis = go_max(mx[0] - mn[0], go_max(mx[1] - mn[1], mx[2] - mn[2]));
    int levels = (int)ceilf(log2f(longAxis / res)) + 1; /* octreerenderer.go:229-231 */
    if (levels <= 1) return -1;
    return levels;
}

typedef struct {
    const go_tree *t;
    const go_lattice *lat;
    float *grid;
    int k0, k1, batch;
    int64_t evals;
    int err;
} slab_job;

/* flatrenderer.go:146-182 evalKRange */
static void *slab_worker(void *arg) {
    slab_job *jb = (slab_job *)arg;
    const go_lattice *L = jb->lat;
    int nx = L->n[0], ny = L->n[1];
    size_t sz = (size_t)(nx + 1) * (ny + 1);
    v3 *pb = (v3 *)malloc(sizeof(v3) * jb->batch);
    float *db = (float *)malloc(sizeof(float) * jb->batch);
    if (!pb || !db) { jb->err = -2; free(pb); free(db); return NULL; }
    size_t batchStart = (size_t)jb->k0 * sz;
    int idx = 0;
    for (int k = jb->k0; k < jb->k1 && !jb->err; k++)
        for (int j = 0; j <= ny && !jb->err; j++)
            for (int i = 0; i <= nx; i++) {
                pb[idx].x = L->origin[0] + (float)i * L->res;
                pb[idx].y = L->origin[1] + (float)j * L->res;
                pb[idx].z = L->origin[2] + (float)k * L->res;
                idx++;
                if (idx == jb->batch) {
                    jb->err = eval3(jb->t, jb->t->root, pb, db, idx);
                    if (jb->err) break;
                    memcpy(jb->grid + batchStart, db, sizeof(float) * idx);
                    jb->evals += idx;
                    batchStart += idx;
                    idx = 0;
                }
            }
    if (idx > 0 && !jb->err) {
        jb->err = eval3(jb->t, jb->t->root, pb, db, idx);
        if (!jb->err) { memcpy(jb->grid + batchStart, db, sizeof(float) * idx); jb->evals += idx; }
    }
    free(pb); free(db);
    return NULL;
}

int64_t go_flat_eval_grid(const go_tree *t, const go_lattice *lat, float *grid, int nthreads, int batch) {
    if (batch < 8 || nthreads < 1) return -1; /* flatrenderer.go:41-46 */
    int nk = lat->n[2] + 1;
    int G = nthreads > nk ? nk : nthreads; /* flatrenderer.go:110-114 */
    slab_job *jobs = (slab_job *)calloc(G, sizeof(slab_job));
    pthread_t *th = (pthread_t *)calloc(G, sizeof(pthread_t));
    if (!jobs || !th) { free(jobs); free(th); return -2; }
    for (int g = 0; g < G; g++) {
        jobs[g].t = t; jobs[g].lat = lat; jobs[g].grid = grid; jobs[g].batch = batch;
        jobs[g].k0 = (int)((int64_t)g * nk / G);       /* flatrenderer.go:120-121 */
        jobs[g].k1 = (int)((int64_t)(g + 1) * nk / G);
    }
    if (G == 1) slab_worker(&jobs[0]);
    else {
        for (int g = 0; g < G; g++) pthread_create(&th[g], NULL, slab_worker, &jobs[g]);
        for (int g = 0; g < G; g++) pthread_join(th[g], NULL);
    }
    int64_t evals = 0; int err = 0;
    for (int g = 0; g < G; g++) { evals += jobs[g].evals; if (jobs[g].err) err = jobs[g].err; }
    free(jobs); free(th);
    return err ? err : evals;
}

/* Corner planes [k0, k1) only (same positions as the whole-lattice call: origin + float32(k)*res with the ABSOLUTE plane
 * index), written to `planes` as (k1-k0) x (ny+1) x (nx+1) floats. Lets tests compare single Z-slabs of lattices that are
 * too large to evaluate whole on the CPU. */
int64_t go_flat_eval_planes(const go_tree *t, const go_lattice *lat, int k0, int k1, float *planes, int nthreads, int batch) {
    if (batch < 8 || nthreads < 1 || k0 < 0 || k1 > lat->n[2] + 1 || k0 >= k1) return -1;
    int nk = k1 - k0;
    int G = nthreads > nk ? nk : nthreads;
    slab_job *jobs = (slab_job *)calloc(G, sizeof(slab_job));
    pthread_t *th = (pthread_t *)calloc(G, sizeof(pthread_t));
    if (!jobs || !th) { free(jobs); free(th); return -2; }
    size_t sz = (size_t)(lat->n[0] + 1) * (lat->n[1] + 1);
    for (int g = 0; g < G; g++) {
        jobs[g].t = t; jobs[g].lat = lat; jobs[g].batch = batch;
        jobs[g].grid = planes - (size_t)k0 * sz; /* the worker indexes by absolute plane; only [k0, k1) is touched */
        jobs[g].k0 = k0 + (int)((int64_t)g * nk / G);
        jobs[g].k1 = k0 + (int)((int64_t)(g + 1) * nk / G);
    }
    if (G == 1) slab_worker(&jobs[0]);
    else {
        for (int g = 0; g < G; g++) pthread_create(&th[g], NULL, slab_worker, &jobs[g]);
        for (int g = 0; g < G; g++) pthread_join(th[g], NULL);
    }
    int64_t evals = 0; int err = 0;
    for (int g = 0; g < G; g++) { evals += jobs[g].evals; if (jobs[g].err) err = jobs[g].err; }
    free(jobs); free(th);
    return err ? err : evals;
}

/* marchcubes.go:76-98 */
static inline void mc_interp(const float *p1, const float *p2, float v1, float v2, float x, float *out) {
    const float eps = 1e-12f;
    int c1 = fabsf(x - v1) < eps, c2 = fabsf(x - v2) < eps;
    if (c1 && !c2) { out[0] = p1[0]; out[1] = p1[1]; out[2] = p1[2]; return; }
    if (c2 && !c1) { out[0] = p2[0]; out[1] = p2[1]; out[2] = p2[2]; return; }
    float tt = 0.5f;
    if (!c1 || !c2) tt = (x - v1) / (v2 - v1);
    out[0] = p1[0] + tt * (p2[0] - p1[0]);
    out[1] = p1[1] + tt * (p2[1] - p1[1]);
    out[2] = p1[2] + tt * (p2[2] - p1[2]);
}

/* marchcubes.go:34-73 */
int go_mc_cube(const float p[24], const float v[8], float tri9[45], int *case_index) {
    int index = 0;
    for (int i = 0; i < 8; i++) if (v[i] < 0) index |= 1 << i;
    if (case_index) *case_index = index;
    int edges = go_mc_edges[index];
    if (edges == 0) return 0;
    float pts[12][3];
    for (int i = 0; i < 12; i++)
        if (edges & (1 << i)) {
            int a = go_mc_pairs[2 * i], b = go_mc_pairs[2 * i + 1];
            mc_interp(p + 3 * a, p + 3 * b, v[a], v[b], 0, pts[i]);
        }
    const int8_t *tb = go_mc_tris + 16 * index;
    int nt = 0;
    for (int i = 0; tb[i] >= 0; i += 3) {
        memcpy(tri9 + 9 * nt + 0, pts[tb[i + 2]], 12);
        memcpy(tri9 + 9 * nt + 3, pts[tb[i + 1]], 12);
        memcpy(tri9 + 9 * nt + 6, pts[tb[i + 0]], 12);
        nt++;
    }
    return nt;
}

/* glrender/glrender.go:9 -- note: a different literal from gsdf.go's sqrt3 */
#define GLRENDER_SQRT3 1.73205080757

/* flatrenderer.go:199-250 */
int64_t go_flat_march_slab(const go_lattice *lat, const float *grid, float *tri9, int64_t max_tris, uint8_t *cases,
                           const uint8_t *blockmask, int cz0, int cz1);

int64_t go_flat_march(const go_lattice *lat, const float *grid, float *tri9, int64_t max_tris, uint8_t *cases,
                      const uint8_t *blockmask) {
    return go_flat_march_slab(lat, grid, tri9, max_tris, cases, blockmask, 0, lat->n[2]);
}

/* Same sweep restricted to cell layers cz in [cz0,cz1): what one rank of the Z-slab partition produces (the
 * reference's own split is evalGrid's k-slabs, flatrenderer.go:120-122). grid/cases/blockmask cover the WHOLE lattice. */
int64_t go_flat_march_slab_w(const go_lattice *lat, const float *grid, float *tri9, int64_t max_tris, uint8_t *cases,
                             const uint8_t *blockmask, int mask_w, int cz0, int cz1);
int64_t go_flat_march_slab(const go_lattice *lat, const float *grid, float *tri9, int64_t max_tris, uint8_t *cases,
                           const uint8_t *blockmask, int cz0, int cz1) {
    return go_flat_march_slab_w(lat, grid, tri9, max_tris, cases, blockmask, 4, cz0, cz1);
}
/* mask_w: width in cells of the cubes blockmask describes (4 = level-3 cubes, 2 = level-2 cubes of a plan that ends with
 * level 2, include/gsdf_b200.h); the mask is [ceil(nz/w)][ceil(ny/w)][ceil(nx/w)]. */
int64_t go_flat_march_slab_w(const go_lattice *lat, const float *grid, float *tri9, int64_t max_tris, uint8_t *cases,
                             const uint8_t *blockmask, int mask_w, int cz0, int cz1) {
    int nx = lat->n[0], ny = lat->n[1], nz = lat->n[2];
    size_t sy = (size_t)nx + 1, sz = sy * ((size_t)ny + 1);
    float r = lat->res;
    float cubeDiag = (float)(2 * GLRENDER_SQRT3) * r; /* Go folds 2*sqrt3 as a constant, then float32 multiply */
    if (mask_w < 1) mask_w = 4;
    int nbx = (nx + mask_w - 1) / mask_w, nby = (ny + mask_w - 1) / mask_w;
    int64_t ntri = 0;
    if (cz0 < 0) cz0 = 0;
    if (cz1 > nz) cz1 = nz;
    for (int cz = cz0; cz < cz1; cz++)
        for (int cy = 0; cy < ny; cy++)
            for (int cx = 0; cx < nx; cx++) {
                size_t ci = (size_t)cx + (size_t)nx * ((size_t)cy + (size_t)ny * cz);
                if (cases) cases[ci] = 0;
                if (blockmask && !blockmask[(size_t)(cx / mask_w) + (size_t)nbx * ((size_t)(cy / mask_w) + (size_t)nby * (cz / mask_w))]) continue;
                size_t base = (size_t)cx + (size_t)cy * sy + (size_t)cz * sz;
                if (fabsf(grid[base]) > cubeDiag) continue;
                float v[8] = {grid[base], grid[base + 1], grid[base + 1 + sy], grid[base + sy],
                              grid[base + sz], grid[base + 1 + sz], grid[base + 1 + sy + sz], grid[base + sy + sz]};
                float ox = lat->origin[0] + (float)cx * r, oy = lat->origin[1] + (float)cy * r, oz = lat->origin[2] + (float)cz * r;
                float p[24] = {ox, oy, oz, ox + r, oy, oz, ox + r, oy + r, oz, ox, oy + r, oz,
                               ox, oy, oz + r, ox + r, oy, oz + r, ox + r, oy + r, oz + r, ox, oy + r, oz + r};
                float tmp[45];
                int idx;
                int nt = go_mc_cube(p, v, tmp, &idx);
                if (cases) cases[ci] = (uint8_t)idx;
                for (int k = 0; k < nt; k++) {
                    if (ntri < max_tris && tri9) memcpy(tri9 + 9 * ntri, tmp + 9 * k, 36);
                    ntri++;
                }
            }
    return ntri;
}

/* octreerenderer.go:180-191 (szDistMult = sqrt3/2 with glrender's sqrt3) and :240-284 (centre rule).
 * Level-3 cubes are 4 cells wide (ms3.Octree.CubeSize = res * 2^(level-1), UNPINNED helper). */
int64_t go_octree_prune_mask(const go_tree *t, const go_lattice *lat, uint8_t *mask) {
    int nbx = (lat->n[0] + 3) / 4, nby = (lat->n[1] + 3) / 4, nbz = (lat->n[2] + 3) / 4;
    size_t nb = (size_t)nbx * nby * nbz;
    v3 *c = (v3 *)malloc(sizeof(v3) * nb);
    float *d = (float *)malloc(sizeof(float) * nb);
    if (!c || !d) { free(c); free(d); return -2; }
    float size = lat->res * 4.0f;
    float half = size * 0.5f;
    size_t i = 0;
    for (int bz = 0; bz < nbz; bz++)
        for (int by = 0; by < nby; by++)
            for (int bx = 0; bx < nbx; bx++, i++) {
                /* CubeCenter = CubeOrigin + size/2 ; CubeOrigin = Origin + res * vec */
                c[i].x = (lat->origin[0] + (float)(4 * bx) * lat->res) + half;
                c[i].y = (lat->origin[1] + (float)(4 * by) * lat->res) + half;
                c[i].z = (lat->origin[2] + (float)(4 * bz) * lat->res) + half;
            }
    int err = eval3(t, t->root, c, d, nb);
    int64_t kept = 0;
    if (!err) {
        float maxDist = size * (float)(GLRENDER_SQRT3 / 2);
        for (i = 0; i < nb; i++) { mask[i] = !(fabsf(d[i]) >= maxDist); kept += mask[i]; }
    }
    free(c); free(d);
    return err ? err : kept;
}

/* Coarse-to-fine prune plan (include/gsdf_b200.h, gsdf_prune_plan): the same rule applied level by level, coarse to fine,
 * to the children of surviving cubes only, with a margin on the threshold: keep iff |d(centre)| < margin * size * sqrt3/2.
 * margin 1 on a single level 3 is go_octree_prune_mask above. Cubes of level L are 2^(L-1) cells wide and aligned to the
 * lattice origin (ms3.Octree.CubeOrigin). mask = level-3 verdicts [nbz][nby][nbx]; *evals = centres evaluated. */
int64_t go_octree_prune_plan(const go_tree *t, const go_lattice *lat, int nlevels, const int *levels, const float *margins, uint8_t *mask,
                             int64_t *evals) {
    /* the plan ends with level 3, or with level 3 followed by level 2 (the mask then describes the 2-cell cubes) */
    if (nlevels < 1) return -1;
    if (levels[nlevels - 1] != 3 && !(levels[nlevels - 1] == 2 && nlevels >= 2 && levels[nlevels - 2] == 3)) return -1;
    uint8_t *parent = NULL;
    int pw = 0, pnx = 0, pny = 0;
    int64_t nev = 0, kept = 0;
    for (int li = 0; li < nlevels; li++) {
        int w = 1 << (levels[li] - 1);
        int ncx = (lat->n[0] + w - 1) / w, ncy = (lat->n[1] + w - 1) / w, ncz = (lat->n[2] + w - 1) / w;
        size_t nc = (size_t)ncx * ncy * ncz;
        uint8_t *cur = li == nlevels - 1 ? mask : (uint8_t *)malloc(nc);
        v3 *c = (v3 *)malloc(sizeof(v3) * nc);
        float *d = (float *)malloc(sizeof(float) * nc);
        size_t *idx = (size_t *)malloc(sizeof(size_t) * nc);
        if (!cur || !c || !d || !idx) { free(c); free(d); free(idx); return -2; }
        float size = lat->res * (float)w;
        float half = size * 0.5f;
        float maxDist = size * (float)(GLRENDER_SQRT3 / 2) * margins[li];
        size_t m = 0, i = 0;
        for (int cz = 0; cz < ncz; cz++)
            for (int cy = 0; cy < ncy; cy++)
                for (int cx = 0; cx < ncx; cx++, i++) {
                    cur[i] = 0;
                    if (parent) {
                        int r = pw / w;
                        if (!parent[((size_t)(cz / r) * pny + (cy / r)) * pnx + (cx / r)]) continue;
                    }
                    c[m].x = (lat->origin[0] + (float)(w * cx) * lat->res) + half;
                    c[m].y = (lat->origin[1] + (float)(w * cy) * lat->res) + half;
                    c[m].z = (lat->origin[2] + (float)(w * cz) * lat->res) + half;
                    idx[m++] = i;
                }
        int err = m ? eval3(t, t->root, c, d, m) : 0;
        if (!err) {
            kept = 0;
            for (size_t k = 0; k < m; k++) { cur[idx[k]] = !(fabsf(d[k]) >= maxDist); kept += cur[idx[k]]; }
            nev += (int64_t)m;
        }
        free(c); free(d); free(idx);
        if (parent) free(parent);
        parent = li == nlevels - 1 ? NULL : cur;
        pw = w; pnx = ncx; pny = ncy;
        if (err) { if (parent) free(parent); return err; }
    }
    if (evals) *evals = nev;
    return kept;
}

/* ------------------------------------------------------------------------------------------------------------
 * STL (glrender/stl.go)
 * ---------------------------------------------------------------------------------------------------------- */
static inline void put_f32(uint8_t *b, float f) { memcpy(b, &f, 4); } /* little-endian host assumed (x86-64) */

int64_t go_stl_write(const float *tri9, int64_t ntri, uint8_t *dst) {
    if (ntri <= 0) return -1;                 /* stl.go:16-18 */
    if (ntri > 0xffffffffLL) return -1;       /* stl.go:21-23 */
    memset(dst, 0, 84);
    uint32_t cnt = (uint32_t)ntri;
    memcpy(dst + 80, &cnt, 4);
    uint8_t *o = dst + 84;
    for (int64_t i = 0; i < ntri; i++, o += 50) {
        const float *t = tri9 + 9 * i;
        /* ms3.Triangle.Normal (gonum r3 style, UNPINNED): Cross(t1-t0, t2-t1); ms3.Unit: Scale(1/Norm(n), n) */
        float s1x = t[3] - t[0], s1y = t[4] - t[1], s1z = t[5] - t[2];
        float s2x = t[6] - t[3], s2y = t[7] - t[4], s2z = t[8] - t[5];
        float nx = s1y * s2z - s1z * s2y, ny = s1z * s2x - s1x * s2z, nz = s1x * s2y - s1y * s2x;
        float inv = 1 / norm3(nx, ny, nz);
        if (nx == 0 && ny == 0 && nz == 0) { nx = ny = nz = NAN; } else { nx *= inv; ny *= inv; nz *= inv; }
        put_f32(o, nx); put_f32(o + 4, ny); put_f32(o + 8, nz);
        memcpy(o + 12, t, 36);
        o[48] = 0; o[49] = 0;
    }
    return 84 + 50 * ntri;
}

int64_t go_stl_read(const uint8_t *src, size_t nbytes, float *tri9, int64_t max_tris) {
    if (nbytes < 84) return -1;
    uint32_t cnt;
    memcpy(&cnt, src + 80, 4);
    if (cnt == 0) return -1; /* stl.go:183-185 */
    if (nbytes < 84 + (size_t)cnt * 50) return -1;
    if (tri9)
        for (int64_t i = 0; i < (int64_t)cnt && i < max_tris; i++) memcpy(tri9 + 9 * i, src + 84 + 50 * i + 12, 36);
    return cnt;
}

/* glrender/image.go:76-105 */
int go_image_eval2(const go_tree *t, const float bbmin[2], const float bbmax[2], int w, int h, float *dist) {
    if (w <= 0 || h <= 0) return -1;
    float dx = (bbmax[0] - bbmin[0]) / (float)w, dy = (bbmax[1] - bbmin[1]) / (float)h;
    float xmin = bbmin[0] + dx / 2;
    v2 *row = (v2 *)malloc(sizeof(v2) * w);
    if (!row) return -2;
    int err = 0;
    for (int j = 0; j < h && !err; j++) {
        float y = bbmax[1] - (float)j * dy; /* un-shifted Max, image.go:92 */
        for (int i = 0; i < w; i++) { row[i].x = (float)i * dx + xmin; row[i].y = y; }
        err = eval2(t, t->root, row, dist + (size_t)j * w, w);
    }
    free(row);
    return err;
}

/* ---- colour conversions: glrender/image.go:50-61, gsdfaux/color.go ------------------------------------------------
 * ms1.SmoothStep / ms1.Interp / ms3.InterpElem are from the un-vendored soypat/geometry module (GLSL smoothstep / mix
 * semantics): parity UNPINNED like the rest of that module. uint8(f) follows Go/amd64: truncate, keep the low byte. */
static uint32_t go_u(float f) { return (uint32_t)(int32_t)f; }
static uint32_t go_rgba(uint32_t r, uint32_t g, uint32_t b, uint32_t a) { return (r & 255u) | (g & 255u) << 8 | (b & 255u) << 16 | (a & 255u) << 24; }
void go_rgb_to_hsv(float r, float g, float b, float hsv[3]) {
    float xmax = r > g ? (r > b ? r : b) : (g > b ? g : b), xmin = r < g ? (r < b ? r : b) : (g < b ? g : b);
    float c = xmax - xmin, h = 0, s = 0, v = xmax;
    if (c == 0) h = 0;
    else if (v == r) h = (g - b) / (c * 6);
    else if (v == g) h = (float)(1.0 / 3) + (b - r) / (c * 6);
    else if (v == b) h = (float)(2.0 / 3) + (r - g) / (c * 6);
    if (h < 0) h += 1;
    if (xmax > 0) s = c / xmax;
    hsv[0] = h; hsv[1] = s; hsv[2] = v;
}
static void go_hsv_to_rgb(float h, float s, float v, float *r, float *g, float *b) { /* color.go:165-188 */
    float c = s * v;
    float x = c * (1 - fabsf(fmodf(h * 6, 2) - 1));
    float m = v - c;
    *r = *g = *b = 0;
    if (h >= 0 && h <= (float)(1.0 / 6)) { *r = c; *g = x; *b = 0; }
    else if (h > (float)(1.0 / 6) && h <= (float)(2.0 / 6)) { *r = x; *g = c; *b = 0; }
    else if (h > (float)(2.0 / 6) && h <= (float)(3.0 / 6)) { *r = 0; *g = c; *b = x; }
    else if (h > (float)(3.0 / 6) && h <= (float)(4.0 / 6)) { *r = 0; *g = x; *b = c; }
    else if (h > (float)(4.0 / 6) && h <= (float)(5.0 / 6)) { *r = x; *g = 0; *b = c; }
    else if (h > (float)(5.0 / 6) && h <= 1.0f) { *r = c; *g = 0; *b = x; }
    *r += m; *g += m; *b += m;
}
uint32_t go_color_of(const go_colorconv *cc, float d) {
    const uint32_t black = 0xff000000u, white = 0xffffffffu, red = 0xff0000ffu;
    if (cc->kind == 1) { /* color.go:77-102 */
        float edge = cc->p[0];
        if (edge == 0) return d < 0 ? black : white;
        float blend = d / edge + 0.5f;
        if (blend <= 0) return black;
        if (blend >= 1) return white;
        blend = clampf(blend, 0, 1);
        uint32_t y = go_u(blend * 255);
        return go_rgba(y, y, y, 255);
    }
    if (cc->kind == 2) { /* color.go:21-47 */
        if (isnan(d)) return red;
        d *= cc->p[0];
        float c[3];
        if (d > 0) { c[0] = 0.9f; c[1] = 0.6f; c[2] = 0.3f; } else { c[0] = 0.65f; c[1] = 0.85f; c[2] = 1.0f; }
        float f = 1 - go_exp(-6 * fabsf(d));
        for (int i = 0; i < 3; i++) c[i] *= f;
        f = 0.8f + 0.2f * go_cos(150 * d);
        for (int i = 0; i < 3; i++) c[i] *= f;
        float t = clampf((fabsf(d) - 0.f) / (0.01f - 0.f), 0, 1);
        t = t * t * (3 - 2 * t);
        float mx = 1 - t;
        for (int i = 0; i < 3; i++) c[i] = c[i] * (1 - mx) + 1.f * mx;
        return go_rgba(go_u(c[0] * 255), go_u(c[1] * 255), go_u(c[2] * 255), 255);
    }
    if (cc->kind == 3) { /* color.go:57-72 */
        float blend = d / cc->p[6] + 0.5f;
        if (blend <= 0) return cc->c0;
        if (blend >= 1) return cc->c1;
        float h0 = cc->p[0], h1 = cc->p[3];
        if (h1 - h0 > 0.5f) h0 += 1.0f;
        else if (h1 - h0 < -0.5f) h1 += 1.0f;
        float h = h0 * (1 - blend) + h1 * blend;
        float s_ = cc->p[1] * (1 - blend) + cc->p[4] * blend;
        float v = cc->p[2] * (1 - blend) + cc->p[5] * blend;
        float r, g, b;
        go_hsv_to_rgb(h, s_, v, &r, &g, &b);
        return go_rgba(go_u(clampf(r, 0, 1) * 255), go_u(clampf(g, 0, 1) * 255), go_u(clampf(b, 0, 1) * 255), 255);
    }
    if (isnan(d) || isinf(d)) return red; /* image.go:50-61 */
    return d > 0 ? white : black;
}
int go_image_render2(const go_tree *t, const float bbmin[2], const float bbmax[2], int w, int h, const go_colorconv *cc, uint8_t *rgba) {
    float *dist = (float *)malloc(sizeof(float) * (size_t)w * h);
    if (!dist) return -2;
    int err = go_image_eval2(t, bbmin, bbmax, w, h, dist);
    go_colorconv def;
    memset(&def, 0, sizeof def);
    if (!cc) cc = &def;
    for (size_t i = 0; !err && i < (size_t)w * h; i++) { uint32_t c = go_color_of(cc, dist[i]); memcpy(rgba + 4 * i, &c, 4); }
    free(dist);
    return err;
}

const int *go_mc_edge_table(void) { return go_mc_edges; }
const int8_t *go_mc_tri_table(void) { return go_mc_tris; }
const int *go_mc_pair_table(void) { return go_mc_pairs; }
