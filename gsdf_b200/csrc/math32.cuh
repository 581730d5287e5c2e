// math32.cuh -- float32 elementary functions for the sm_100a kernels and the host-side flattener.
//
// The reference computes with github.com/chewxy/math32 v1.11.1 (go.mod:8), a float32 port of Go's math package
// (Cephes-derived rational/polynomial kernels, three-part pi/4 argument reduction).  CUDA's sinf/cosf/atan2f are
// different algorithms and would flip the sign of a distance a few ulp from zero now and then, which changes
// marching-cubes case indices.  So the kernels evaluate the same published algorithms with the same operation
// order, individually rounded (compile with -fmad=false; IEEE div/sqrt are the nvcc defaults without fast-math).
//
// Everything is __host__ __device__ so the flattener pre-computes constants (tan(taper), sincos of fixed angles)
// with the same code the kernels run.
#pragma once
#ifdef __CUDACC_RTC__
#include "rtc_types.cuh"
#else
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#endif

#define M32_HD __host__ __device__ __forceinline__
// Large helpers (hypot, atan2, sincos) can be kept out of line to shrink the interpreter's instruction footprint
// (the evaluate kernel stalls on instruction fetch when every opcode body inlines its own copies).
#ifdef GSDF_NOINLINE_MATH
#define M32_BIG static __host__ __device__ __noinline__
#else
#define M32_BIG __host__ __device__ __forceinline__
#endif

namespace m32 {

constexpr double kPi = 3.14159265358979323846264338327950288419716939937510582097494459;
constexpr float kPiF = (float)kPi;
constexpr float kTwoPiF = (float)(2 * kPi);

M32_HD float absf(float x) { return fabsf(x); }
// math32.Min/Max (Go math semantics): a NaN operand gives NaN, -0 orders below +0. On the device that is exactly PTX
// min.NaN.f32 / max.NaN.f32 (sm_80+), one instruction like fminf/fmaxf -- which would DROP the NaN and let a union or
// difference turn an invalid operand (ellipse2D far outside its bounds) into a finite distance. One corner is left:
// Go returns -Inf for Min(NaN, -Inf) (+Inf for Max(NaN, +Inf)) because it tests the infinity first; here that is NaN.
M32_HD float minf(float a, float b) {
#ifdef __CUDA_ARCH__
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
#else
    if (a != a || b != b) return __builtin_nanf("");
    if (a == 0.f && b == 0.f) return __builtin_signbit(a) ? a : b;
    return a < b ? a : b;
#endif
}
M32_HD float maxf(float a, float b) {
#ifdef __CUDA_ARCH__
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
#else
    if (a != a || b != b) return __builtin_nanf("");
    if (a == 0.f && b == 0.f) return __builtin_signbit(a) ? b : a;
    return a > b ? a : b;
#endif
}

M32_HD float sqrt(float x) {
#ifdef __CUDA_ARCH__
    return __fsqrt_rn(x);
#else
    return sqrtf(x);
#endif
}
M32_HD float div(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
M32_HD float mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
M32_HD float add(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}

// math32.Hypot: p*Sqrt(1+(q/p)^2) with p>=q (NOT sqrt(p*p+q*q)).
M32_BIG float hypot32(float p, float q) {
    p = fabsf(p);
    q = fabsf(q);
    if (p < q) { float t = p; p = q; q = t; }
    if (p == 0.0f) return 0.0f;
    q = div(q, p);
    return mul(p, sqrt(add(1.0f, mul(q, q))));
}
// soypat/geometry ms3.Norm / ms2.Norm (gonum r3/r2 style): nested Hypot.
M32_HD float norm3(float x, float y, float z) { return hypot32(x, hypot32(y, z)); }
M32_HD float norm2(float x, float y) { return hypot32(x, y); }

// gsdf.go:148-167
M32_HD float signf(float a) { return a == 0.0f ? 0.0f : copysignf(1.0f, a); }
M32_HD float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
M32_HD float mixf(float x, float y, float a) { return add(mul(x, add(1.0f, -a)), mul(y, a)); }

// ---- atan / atan2 (math32 atan.go, atan2.go) ----
M32_HD float xatan(float x) {
    const float P0 = -8.750608600031904122785e-01f, P1 = -1.615753718733365076637e+01f,
                P2 = -7.500855792314704667340e+01f, P3 = -1.228866684490136173410e+02f,
                P4 = -6.485021904942025371773e+01f;
    const float Q0 = +2.485846490142306297962e+01f, Q1 = +1.650270098316988542046e+02f,
                Q2 = +4.328810604912902668951e+02f, Q3 = +4.853903996359136964868e+02f,
                Q4 = +1.945506571482613964425e+02f;
    float z = mul(x, x);
    float num = add(mul(add(mul(add(mul(add(mul(P0, z), P1), z), P2), z), P3), z), P4);
    float den = add(mul(add(mul(add(mul(add(mul(add(z, Q0), z), Q1), z), Q2), z), Q3), z), Q4);
    z = div(mul(z, num), den);
    return add(mul(x, z), x);
}
// atan.go's satan has three ranges, each with its own division and xatan call; atan mirrors negative arguments. Here the
// OPERANDS of the one division and the terms of the final sum are selected by range and every lane runs the same
// instructions: a warp whose lanes fall into different ranges (any warp that sees a few degrees of a screw) no longer
// executes all three paths one after the other, and the code is a third of the size. Same operations on the same values
// as the three-path form, so the same bits: x / 1 == x exactly, and atan(-x) evaluates satan(|x|) as before.
M32_HD float satan(float x) {
    const float Morebits = 6.123233995736765886130e-17f;
    const float Tan3pio8 = 2.41421356237309504880f;
    const bool small = x <= 0.66f;
    const bool big = x > Tan3pio8;
    const float num = small ? x : (big ? 1.0f : add(x, -1.0f));
    const float den = small ? 1.0f : (big ? x : add(x, 1.0f));
    const float t = xatan(div(num, den));
    const float r = add(add(big ? (float)(kPi / 2) : (float)(kPi / 4), big ? -t : t), big ? Morebits : mul(0.5f, Morebits));
    return small ? t : r;
}
M32_HD float atan(float x) {
    const float s = satan(fabsf(x));
    return x == 0.0f ? x : (x > 0.0f ? s : -s);
}
M32_BIG float atan2(float y, float x) {
    // atan2.go tests its special operands (NaN, zero, infinity) one by one in front of the quotient; lattice points are
    // finite and non-zero, so ONE test sends them straight to the quotient and the chain stays out of their way
    const float ay = fabsf(y), ax = fabsf(x);
    if (ay > 0.0f && ay < INFINITY && ax > 0.0f && ax < INFINITY) {
        float q = atan(div(y, x));
        if (x < 0.0f) return q <= 0.0f ? add(q, kPiF) : add(q, -kPiF);
        return q;
    }
    if (y != y || x != x) return NAN;
    if (y == 0.0f) {
        if (x >= 0.0f && !signbit(x)) return copysignf(0.0f, y);
        return copysignf(kPiF, y);
    }
    if (x == 0.0f) return copysignf((float)(kPi / 2), y);
    if (isinf(x)) {
        if (x > 0.0f) return isinf(y) ? copysignf((float)(kPi / 4), y) : copysignf(0.0f, y);
        return isinf(y) ? copysignf((float)(3 * kPi / 4), y) : copysignf(kPiF, y);
    }
    return copysignf((float)(kPi / 2), y);  // y is infinite, x finite and non-zero
}

// ---- sin / cos / tan (math32 sin.go, tan.go) ----
constexpr float kPI4A = 7.85398125648498535156e-1f;
constexpr float kPI4B = 3.77489470793079817668e-8f;
constexpr float kPI4C = 2.69515142907905952645e-15f;
constexpr float kM4PI = 1.273239544735162542821171882678754627704620361328125f;

M32_HD float sin_poly(float z, float zz) {
    const float S0 = 1.58962301576546568060e-10f, S1 = -2.50507477628578072866e-8f, S2 = 2.75573136213857245213e-6f,
                S3 = -1.98412698295895385996e-4f, S4 = 8.33333333332211858878e-3f, S5 = -1.66666666666666307295e-1f;
    float p = add(mul(add(mul(add(mul(add(mul(add(mul(S0, zz), S1), zz), S2), zz), S3), zz), S4), zz), S5);
    return add(z, mul(mul(z, zz), p));
}
M32_HD float cos_poly(float zz) {
    const float C0 = -1.13585365213876817300e-11f, C1 = 2.08757008419747316778e-9f, C2 = -2.75573141792967388112e-7f,
                C3 = 2.48015872888517045348e-5f, C4 = -1.38888888888730564116e-3f, C5 = 4.16666666666665929218e-2f;
    float p = add(mul(add(mul(add(mul(add(mul(add(mul(C0, zz), C1), zz), C2), zz), C3), zz), C4), zz), C5);
    return add(add(1.0f, -mul(0.5f, zz)), mul(mul(zz, zz), p));
}
M32_HD float trig_reduce(float x, uint32_t &j) {
    uint32_t jj = (uint32_t)mul(x, kM4PI);
    float y = (float)jj;
    if (jj & 1u) { jj++; y = add(y, 1.0f); }
    j = jj & 7u;
    return add(add(add(x, -mul(y, kPI4A)), -mul(y, kPI4B)), -mul(y, kPI4C));
}
M32_HD float sin(float x) {
    if (x == 0.0f || x != x) return x;
    if (isinf(x)) return NAN;
    bool sign = false;
    if (x < 0.0f) { x = -x; sign = true; }
    uint32_t j;
    float z = trig_reduce(x, j);
    if (j > 3u) { sign = !sign; j -= 4u; }
    float zz = mul(z, z);
    float y = (j == 1u || j == 2u) ? cos_poly(zz) : sin_poly(z, zz);
    return sign ? -y : y;
}
M32_HD float cos(float x) {
    if (x != x || isinf(x)) return NAN;
    bool sign = false;
    x = fabsf(x);
    uint32_t j;
    float z = trig_reduce(x, j);
    if (j > 3u) { j -= 4u; sign = !sign; }
    if (j > 1u) sign = !sign;
    float zz = mul(z, z);
    float y = (j == 1u || j == 2u) ? sin_poly(z, zz) : cos_poly(zz);
    return sign ? -y : y;
}
// math32.Sincos shares one reduction; the two results equal Sin(x), Cos(x) bit for bit.
M32_BIG void sincos(float x, float &s, float &c) {
    if (x == 0.0f) { s = x; c = 1.0f; return; }
    if (x != x || isinf(x)) { s = NAN; c = NAN; return; }
    bool ssign = false, csign = false;
    if (x < 0.0f) { x = -x; ssign = true; }
    uint32_t j;
    float z = trig_reduce(x, j);
    if (j > 3u) { j -= 4u; ssign = !ssign; csign = !csign; }
    if (j > 1u) csign = !csign;
    float zz = mul(z, z);
    float cp = cos_poly(zz), sp = sin_poly(z, zz);
    if (j == 1u || j == 2u) { s = cp; c = sp; } else { s = sp; c = cp; }
    if (ssign) s = -s;
    if (csign) c = -c;
}
M32_HD float tan(float x) {
    const float P0 = -1.30936939181383777646e4f, P1 = 1.15351664838587416140e6f, P2 = -1.79565251976484877988e7f;
    const float Q1 = 1.36812963470692954678e4f, Q2 = -1.32089234440210967447e6f, Q3 = 2.50083801823357915839e7f,
                Q4 = -5.38695755929454629881e7f;
    if (x == 0.0f || x != x) return x;
    if (isinf(x)) return NAN;
    bool sign = false;
    if (x < 0.0f) { x = -x; sign = true; }
    uint32_t j = (uint32_t)mul(x, kM4PI);
    float y = (float)j;
    if (j & 1u) { j++; y = add(y, 1.0f); }
    float z = add(add(add(x, -mul(y, kPI4A)), -mul(y, kPI4B)), -mul(y, kPI4C));
    float zz = mul(z, z);
    if (zz > 1e-14f) {
        float num = mul(zz, add(mul(add(mul(P0, zz), P1), zz), P2));
        float den = add(mul(add(mul(add(mul(add(zz, Q1), zz), Q2), zz), Q3), zz), Q4);
        y = add(z, mul(z, div(num, den)));
    } else {
        y = z;
    }
    if (j & 2u) y = div(-1.0f, y);
    return sign ? -y : y;
}

// math32.Asin / Acos (asin.go): used only host-side by PolygonBuilder.Smooth.
M32_HD float asin(float x) {
    if (x == 0.0f) return x;
    bool sign = false;
    if (x < 0.0f) { x = -x; sign = true; }
    if (x > 1.0f) return NAN;
    float temp = sqrt(add(1.0f, -mul(x, x)));
    if (x > 0.7f) temp = add((float)(kPi / 2), -satan(div(temp, x)));
    else temp = satan(div(x, temp));
    return sign ? -temp : temp;
}
M32_HD float acos(float x) { return add((float)(kPi / 2), -asin(x)); }

// math32.Cbrt restated as FreeBSD's cbrtf: integer seed, two Newton steps in double precision, one rounding to float32.
// Double arithmetic uses explicit round-to-nearest intrinsics on the device so nothing is contracted.
M32_HD float cbrt32(float x) {
    uint32_t hx;
#ifdef __CUDA_ARCH__
    hx = __float_as_uint(x);
#else
    memcpy(&hx, &x, 4);
#endif
    const uint32_t sign = hx & 0x80000000u;
    hx ^= sign;
    if (hx >= 0x7f800000u) return add(x, x);
    float t;
    if (hx < 0x00800000u) {
        if (hx == 0u) return x;
        uint32_t w = 0x4b800000u;  // 2**24
#ifdef __CUDA_ARCH__
        t = mul(__uint_as_float(w), x);
        w = sign | ((__float_as_uint(t) & 0x7fffffffu) / 3u + 642849266u);
        t = __uint_as_float(w);
#else
        memcpy(&t, &w, 4); t *= x; memcpy(&w, &t, 4);
        w = sign | ((w & 0x7fffffffu) / 3u + 642849266u);
        memcpy(&t, &w, 4);
#endif
    } else {
        const uint32_t w = sign | (hx / 3u + 709958130u);
#ifdef __CUDA_ARCH__
        t = __uint_as_float(w);
#else
        memcpy(&t, &w, 4);
#endif
    }
#ifdef __CUDA_ARCH__
    const double xd = (double)x;
    double T = (double)t, r = __dmul_rn(__dmul_rn(T, T), T);
    T = __ddiv_rn(__dmul_rn(T, __dadd_rn(__dadd_rn(xd, xd), r)), __dadd_rn(__dadd_rn(xd, r), r));
    r = __dmul_rn(__dmul_rn(T, T), T);
    T = __ddiv_rn(__dmul_rn(T, __dadd_rn(__dadd_rn(xd, xd), r)), __dadd_rn(__dadd_rn(xd, r), r));
    return (float)T;
#else
    double T = t, r = T * T * T;
    T = T * ((double)x + x + r) / (x + r + r);
    r = T * T * T;
    T = T * ((double)x + x + r) / (x + r + r);
    return (float)T;
#endif
}

// math32 log.go / exp.go (ports of Go's math.Log / math.Exp), float32 arithmetic.
M32_HD float log32(float x) {
    const float Ln2Hi = 6.93147180369123816490e-01f, Ln2Lo = 1.90821492927058770002e-10f;
    const float L1 = 6.666666666666735130e-01f, L2 = 3.999999999940941908e-01f, L3 = 2.857142874366239149e-01f,
                L4 = 2.222219843214978396e-01f, L5 = 1.818357216161805012e-01f, L6 = 1.531383769920937332e-01f,
                L7 = 1.479819860511658591e-01f;
    if (x != x || (isinf(x) && x > 0.f)) return x;
    if (x < 0.f) return NAN;
    if (x == 0.f) return -INFINITY;
    int ki;
    float f1 = frexpf(x, &ki);
    if (f1 < (float)(1.41421356237309504880168872420969808 / 2)) { f1 = mul(f1, 2.f); ki--; }
    const float f = add(f1, -1.f);
    const float k = (float)ki;
    const float s_ = div(f, add(2.f, f));
    const float s2 = mul(s_, s_);
    const float s4 = mul(s2, s2);
    const float t1 = mul(s2, add(L1, mul(s4, add(L3, mul(s4, add(L5, mul(s4, L7)))))));
    const float t2 = mul(s4, add(L2, mul(s4, add(L4, mul(s4, L6)))));
    const float R = add(t1, t2);
    const float hfsq = mul(mul(0.5f, f), f);
    return add(mul(k, Ln2Hi), -add(add(hfsq, -add(mul(s_, add(hfsq, R)), mul(k, Ln2Lo))), -f));
}
M32_HD float exp32(float x) {
    const float Ln2Hi = 6.93147180369123816490e-01f, Ln2Lo = 1.90821492927058770002e-10f, Log2e = 1.44269504088896338700e+00f;
    const float P1 = 1.66666666666666657415e-01f, P2 = -2.77777777770155933842e-03f, P3 = 6.61375632143793436117e-05f,
                P4 = -1.65339022054652515390e-06f, P5 = 4.13813679705723846039e-08f;
    if (x != x || (isinf(x) && x > 0.f)) return x;
    if (isinf(x)) return 0.f;
    if (x > 88.72283905206835f) return INFINITY;
    if (x < -103.97207708f) return 0.f;
    int k = 0;
    if (x < 0.f) k = (int)add(mul(Log2e, x), -0.5f);
    else if (x > 0.f) k = (int)add(mul(Log2e, x), 0.5f);
    const float hi = add(x, -mul((float)k, Ln2Hi));
    const float lo = mul((float)k, Ln2Lo);
    const float r = add(hi, -lo);
    const float t = mul(r, r);
    const float c = add(r, -mul(t, add(P1, mul(t, add(P2, mul(t, add(P3, mul(t, add(P4, mul(t, P5))))))))));
    const float y = add(1.f, -add(add(lo, -div(mul(r, c), add(2.f, -c))), -hi));
    return ldexpf(y, k);
}
// math32.Pow(x, y) for x >= 0, 0 < y < 0.5: Exp(y*Log(x)) (Go pow.go with yi = 0).
M32_HD float pow_frac(float x, float y) {
    if (x == 0.f) return 0.f;
    if (x == 1.f) return 1.f;
    return exp32(mul(y, log32(x)));
}

}  // namespace m32
