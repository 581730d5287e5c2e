// interp.cuh -- the node-program interpreter: one stack machine per point, P points per thread.
//
// Every thread of a CTA walks the SAME instruction stream held in shared memory, so opcode dispatch is warp-uniform
// (one indirect branch per instruction, no divergence); only a few primitives branch per point inside their bodies.
// Per-point formulas restate cpu_evaluators.go / forge/threads/threads.go:141-202 operation by operation; the file
// must be compiled with -fmad=false so each float32 op rounds individually, as Go/amd64 does.
//
// Stacks live in shared memory, laid out [slot][component][point][thread] so a warp access is one conflict-free
// 128-byte wavefront. The distance-stack top and the current position stay in registers.
#pragma once
#ifdef __CUDACC_RTC__
#include "rtc_types.cuh"
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "../../include/gsdf_program.h"
#include "math32.cuh"

// Lockstep: a CTA barrier per interpreted instruction keeps all warps of the CTA inside the same opcode body, so the
// CTA is ONE instruction-cache stream (see kernels.cuh for the measured effect). Define GSDF_NO_LOCKSTEP to disable.
#if !defined(GSDF_NO_LOCKSTEP) && !defined(GSDF_LOCKSTEP)
#define GSDF_LOCKSTEP 1
#endif

// The vote of a slab / box guard (gsdf_program.h): all points of the voting unit must agree before the unit jumps. The
// interpreter votes CTA-wide (its warps walk the program in lockstep anyway); the run-time compiled kernels (jit.cu), which
// have no lockstep barriers, define GSDF_WARP_GUARDS and vote per warp -- a guard is value-preserving at any granularity,
// so results do not depend on the choice.
#ifdef GSDF_WARP_GUARDS
#define GSDF_GUARD_VOTE(pred) __all_sync(0xffffffffu, (pred))
#else
#define GSDF_GUARD_VOTE(pred) __syncthreads_and(pred)
#endif

namespace gsdfk {

template <int P>
struct Machine {
    float px[P], py[P], pz[P];  // current position
    float top[P];               // distance stack top
    float *dstk;                // &dstack[threadIdx.x]
    float *pstk;                // &pstack[threadIdx.x]
    int stride;                 // blockDim.x
    int dsp, psp;
    bool skip;                  // CTA-uniform: a slab guard fired, the next combiner keeps `a` (gsdf_program.h)
#ifdef GSDF_RXY
    float *rxy;                 // &radius cache[threadIdx.x] (experimental radius reuse, gsdf_program.h)
    // r = Hypot(px, py), from the cache when the flattener proved it holds the radius of bit-identical x, y
    __device__ __forceinline__ void radius(uint32_t flags, float (&r)[P]) {
        if (flags & GSDF_RXY_READ) {
#pragma unroll
            for (int j = 0; j < P; j++) r[j] = rxy[j * stride];
        } else {
#pragma unroll
            for (int j = 0; j < P; j++) r[j] = m32::hypot32(px[j], py[j]);
            if (flags & GSDF_RXY_WRITE) {
#pragma unroll
                for (int j = 0; j < P; j++) rxy[j * stride] = r[j];
            }
        }
    }
#endif

    __device__ __forceinline__ void init(float *d, float *p, int s) {
        dstk = d; pstk = p; stride = s; dsp = -1; psp = 0; skip = false;
#pragma unroll
        for (int j = 0; j < P; j++) top[j] = 0.f;
    }
    __device__ __forceinline__ void pushD() {
        int s = dsp < 0 ? 0 : dsp;
#pragma unroll
        for (int j = 0; j < P; j++) dstk[(s * P + j) * stride] = top[j];
        dsp++;
    }
    __device__ __forceinline__ void popBelow(float (&a)[P]) {
        dsp--;
#pragma unroll
        for (int j = 0; j < P; j++) a[j] = dstk[(dsp * P + j) * stride];
    }
    __device__ __forceinline__ void pushPos(const float (&x)[P], const float (&y)[P], const float (&z)[P]) {
#pragma unroll
        for (int j = 0; j < P; j++) {
            pstk[((psp * 3 + 0) * P + j) * stride] = x[j];
            pstk[((psp * 3 + 1) * P + j) * stride] = y[j];
            pstk[((psp * 3 + 2) * P + j) * stride] = z[j];
        }
        psp++;
    }
    __device__ __forceinline__ void loadPos(int slot) {
#pragma unroll
        for (int j = 0; j < P; j++) {
            px[j] = pstk[((slot * 3 + 0) * P + j) * stride];
            py[j] = pstk[((slot * 3 + 1) * P + j) * stride];
            pz[j] = pstk[((slot * 3 + 2) * P + j) * stride];
        }
    }
};

__device__ __forceinline__ float4 ldf4(const uint4 *prog, int i) {
    uint4 u = prog[i];
    return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
}

// Slab-guard predicate (include/gsdf_program.h): true when a subtree whose value is >= w cannot change the result of
// the combiner that follows, given the value `a` already on top of the distance stack. Strict comparisons (and a != 0
// for the smooth blend) keep the skipped result bit-identical, signed zeros included.
__device__ __forceinline__ bool guard_dead(uint32_t kind, float w, float a, float k) {
    if (kind == GSDF_GUARD_DIFF) return -w < a;
    if (kind == GSDF_GUARD_MIN) return w > a;
    return (w - a) >= k && a != 0.f;
}

// EXT selects the instantiation that also carries the rarely used heavy 2-D primitives (ellipse2D, quadbezier2d: double
// precision cbrt, exp/log). Keeping them out of the default kernel matters: with them compiled in, the kernel needs a
// real call stack and the common path slows down by ~20 % (measured); programs that contain them run the EXT kernel.
// One instruction: h = prog[pc] (the caller loads it; a specialised kernel passes the opcode word as a constant, so that the
// switch folds to the one body). Returns false at END; pc moves to the next instruction or to a guard's target.
template <int P, bool EXT>
__device__ __forceinline__ bool exec_one(Machine<P> &m, const uint4 h, const uint4 *__restrict__ prog, int &pc, const float4 *__restrict__ aux) {
    using namespace m32;
    {
        const uint32_t op = h.x & 0xffu;
        const int len = (int)((h.x >> 8) & 0xffu);
        const float f2 = __uint_as_float(h.z), f3 = __uint_as_float(h.w);
        switch (op) {
        case GSDF_OP_END:
            return false;
        // ------------------------------------------------------------------ 3D primitives
        case GSDF_OP_SPHERE: {  // cpu_evaluators.go:20-26
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = norm3(m.px[j], m.py[j], m.pz[j]) - f2;
        } break;
        case GSDF_OP_BOX: {  // :28-36
            const float4 c = ldf4(prog, pc + 1);
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float qx = (absf(m.px[j]) - c.x) + c.w, qy = (absf(m.py[j]) - c.y) + c.w, qz = (absf(m.pz[j]) - c.z) + c.w;
                m.top[j] = norm3(maxf(qx, 0.f), maxf(qy, 0.f), maxf(qz, 0.f)) + minf(maxf(qx, maxf(qy, qz)), 0.f) - c.w;
            }
        } break;
        case GSDF_OP_BOXFRAME: {  // :38-57
            const float4 c = ldf4(prog, pc + 1);
            const float e = c.w;
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = absf(m.px[j]) - c.x, y = absf(m.py[j]) - c.y, z = absf(m.pz[j]) - c.z;
                float qx = absf(x + e) + (-e), qy = absf(y + e) + (-e), qz = absf(z + e) + (-e);
                float s1 = minf(0.f, maxf(x, maxf(qy, qz)));
                float n1 = norm3(maxf(x, 0.f), maxf(qy, 0.f), maxf(qz, 0.f)) + s1;
                float s2 = minf(0.f, maxf(qx, maxf(y, qz)));
                float n2 = norm3(maxf(qx, 0.f), maxf(y, 0.f), maxf(qz, 0.f)) + s2;
                float s3 = minf(0.f, maxf(qx, maxf(qy, z)));
                float n3 = norm3(maxf(qx, 0.f), maxf(qy, 0.f), maxf(z, 0.f)) + s3;
                m.top[j] = minf(n1, minf(n2, n3));
            }
        } break;
        case GSDF_OP_TORUS: {  // :59-68  f2=rGreater f3=rLesser
            m.pushD();
#ifdef GSDF_RXY
            float r[P];
            m.radius(h.y, r);
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = norm2(r[j] - f2, m.pz[j]) - f3;
#else
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = norm2(hypot32(m.px[j], m.py[j]) - f2, m.pz[j]) - f3;
#endif
        } break;
        case GSDF_OP_CYLINDER: {  // :70-88
            const float4 c = ldf4(prog, pc + 1);
            m.pushD();
#ifdef GSDF_RXY
            {
                float r[P];
                m.radius(h.y, r);
                if ((h.y & 1u) == 0u) {
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        float dx = r[j] - c.x;
                        float dy = absf(m.pz[j]) - c.y;
                        m.top[j] = minf(0.f, maxf(dx, dy)) + hypot32(maxf(0.f, dx), maxf(0.f, dy));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        float dx = r[j] - c.x + c.z;
                        float dy = absf(m.pz[j]) - c.y;
                        m.top[j] = minf(maxf(dx, dy), 0.f) + hypot32(maxf(dx, 0.f), maxf(dy, 0.f)) - c.z;
                    }
                }
                break;
            }
#endif
            if (h.y == 0u) {
#pragma unroll
                for (int j = 0; j < P; j++) {
                    float dx = hypot32(m.px[j], m.py[j]) - c.x;
                    float dy = absf(m.pz[j]) - c.y;
                    m.top[j] = minf(0.f, maxf(dx, dy)) + hypot32(maxf(0.f, dx), maxf(0.f, dy));
                }
            } else {
#pragma unroll
                for (int j = 0; j < P; j++) {
                    float dx = hypot32(m.px[j], m.py[j]) - c.x + c.z;
                    float dy = absf(m.pz[j]) - c.y;
                    m.top[j] = minf(maxf(dx, dy), 0.f) + hypot32(maxf(dx, 0.f), maxf(dy, 0.f)) - c.z;
                }
            }
        } break;
        case GSDF_OP_HEX: {  // :90-105  c=(side,h,clm)
            const float4 c = ldf4(prog, pc + 1);
            const float k1 = (float)(-0.8660254037844386467637231707529361834714026269051903140279034897);
            const float twok1 = (float)(2 * -0.8660254037844386467637231707529361834714026269051903140279034897);
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = absf(m.px[j]), y = absf(m.py[j]), z = absf(m.pz[j]);
                float pm = minf(k1 * x + 0.5f * y, 0.f);
                x -= twok1 * pm;
                y -= 1.0f * pm;
                float d1 = hypot32(x - clampf(x, -c.z, c.z), y - c.x) * signf(y - c.x);
                float d2 = z - c.y;
                m.top[j] = minf(maxf(d1, d2), 0.f) + hypot32(maxf(d1, 0.f), maxf(d2, 0.f));
            }
        } break;
        // ------------------------------------------------------------------ 2D primitives
        case GSDF_OP_CIRCLE2D: {  // :661
            m.pushD();
#ifdef GSDF_RXY
            float r[P];
            m.radius(h.y, r);  // ms2.Norm = Hypot(x, y)
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = r[j] - f2;
#else
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = norm2(m.px[j], m.py[j]) - f2;
#endif
        } break;
        case GSDF_OP_RECT2D: {  // :685  f2=bx f3=by
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float dx = absf(m.px[j]) - f2, dy = absf(m.py[j]) - f3;
                m.top[j] = norm2(maxf(dx, 0.f), maxf(dy, 0.f)) + minf(0.f, maxf(dx, dy));
            }
        } break;
        case GSDF_OP_LINE2D: {  // :551  c1=(ax,ay,bax,bay) c2=(dotba,w)
            const float4 c1 = ldf4(prog, pc + 1), c2 = ldf4(prog, pc + 2);
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float pax = m.px[j] - c1.x, pay = m.py[j] - c1.y;
                float hh = clampf((pax * c1.z + pay * c1.w) / c2.x, 0.f, 1.f);
                m.top[j] = norm2(pax - hh * c1.z, pay - hh * c1.w) - c2.y;
            }
        } break;
        case GSDF_OP_LINES2D: {  // :1145  w1=aux_off(floats) w2=nseg w3=w
            const float4 *seg = aux + (h.y >> 2);
            const int nseg = (int)h.z;
            float d[P];
#pragma unroll
            for (int j = 0; j < P; j++) d[j] = 1e23f;
            for (int s = 0; s < nseg; s++) {
                const float4 ab = seg[s];
                const float bax = ab.z - ab.x, bay = ab.w - ab.y;
                const float dotba = bax * bax + bay * bay;
#pragma unroll
                for (int j = 0; j < P; j++) {
                    float pax = m.px[j] - ab.x, pay = m.py[j] - ab.y;
                    float hh = clampf((pax * bax + pay * bay) / dotba, 0.f, 1.f);
                    float ex = pax - hh * bax, ey = pay - hh * bay;
                    d[j] = minf(d[j], ex * ex + ey * ey);
                }
            }
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = m32::sqrt(d[j]) - f3;
        } break;
        case GSDF_OP_ARC2D: {  // :564  c1=(r,t,s,c) c2=(scrx,scry)
            const float4 c1 = ldf4(prog, pc + 1), c2 = ldf4(prog, pc + 2);
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = absf(m.px[j]), y = m.py[j];
                m.top[j] = (c1.w * x > c1.z * y) ? norm2(x - c2.x, y - c2.y) - c1.y : absf(norm2(x, y) - c1.x) - c1.y;
            }
        } break;
        case GSDF_OP_EQTRI2D: {  // :669  f2=r f3=r/k
            const float k = (float)1.7320508075688772935274463415058723669428052538103806280558069794;
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = absf(m.px[j]) - f2, y = m.py[j] + f3;
                if (x + k * y > 0.f) {
                    float nx = x - k * y, ny = -k * x - y;
                    x = 0.5f * nx; y = 0.5f * ny;
                }
                x -= clampf(x, -2.f * f2, 0.f);
                m.top[j] = -norm2(x, y) * signf(y);
            }
        } break;
        case GSDF_OP_HEX2D: {  // :718  f2=r f3=kz*r
            const float kx = (float)(-0.8660254037844386467637231707529361834714026269051903140279034897), ky = 0.5f;
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = absf(m.px[j]), y = absf(m.py[j]);
                float mm = 2.f * minf(kx * x + ky * y, 0.f);
                x = x - mm * kx; y = y - mm * ky;
                x = x - clampf(x, -f3, f3); y = y - f2;
                m.top[j] = signf(y) * norm2(x, y);
            }
        } break;
        case GSDF_OP_OCT2D: {  // :731  f2=r f3=kz*r
            const float kx = -0.9238795325f, ky = 0.3826834323f;
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = absf(m.px[j]), y = absf(m.py[j]);
                float mm = 2.f * minf(kx * x + ky * y, 0.f);
                x = x - mm * kx; y = y - mm * ky;
                mm = 2.f * minf(-kx * x + ky * y, 0.f);
                x = x - mm * -kx; y = y - mm * ky;
                x = x - clampf(x, -f3, f3); y = y - f2;
                m.top[j] = signf(y) * norm2(x, y);
            }
        } break;
        case GSDF_OP_DIAMOND2D: {  // :694  c=(bx,by,dot(b,b))
            const float4 c = ldf4(prog, pc + 1);
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = absf(m.px[j]), y = absf(m.py[j]);
                float ux = c.x - 2.f * x, uy = c.y - 2.f * y;
                float hh = clampf((ux * c.x - uy * c.y) / c.z, -1.f, 1.f);
                float d = norm2(x - (0.5f * c.x) * (1.f - hh), y - (0.5f * c.y) * (1.f + hh));
                m.top[j] = d * signf(x * c.y + y * c.x - c.x * c.y);
            }
        } break;
        case GSDF_OP_ROUNDX2D: {  // :705  f2=w f3=r
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = absf(m.px[j]), y = absf(m.py[j]);
                float sub = 0.5f * minf(x + y, f2);
                m.top[j] = norm2(x - sub, y - sub) - f3;
            }
        } break;
        case GSDF_OP_ELLIPSE2D: if constexpr (EXT) {  // :750-791  f2=a f3=b
            const float sq3 = (float)1.7320508075688772935274463415058723669428052538103806280558069794;
            m.pushD();
#pragma unroll 1
            for (int j = 0; j < P; j++) {
                float a = f2, b = f3;
                float x = absf(m.px[j]), y = absf(m.py[j]);
                if (x > y) { float t = x; x = y; y = t; t = a; a = b; b = t; }
                const float l = b * b - a * a;
                const float mm = a * x / l, m2 = mm * mm;
                const float nn = b * y / l, n2 = nn * nn;
                const float c = (m2 + n2 - 1.f) / 3.f;
                const float c3 = c * c * c;
                const float q = c3 + 2.f * m2 * n2;
                const float d = c3 + m2 * n2;
                const float g = mm + mm * n2;
                float co;
                if (d < 0.f) {
                    const float hh = m32::acos(q / c3) / 3.f;
                    float sh, ch;
                    m32::sincos(hh, sh, ch);
                    const float t = sq3 * sh;
                    const float rx = m32::sqrt(-c * (ch + t + 2.f) + m2);
                    const float ry = m32::sqrt(-c * (ch - t + 2.f) + m2);
                    co = (ry + signf(l) * rx + absf(g) / (rx * ry) - mm) / 2.f;
                } else {
                    const float hh = 2.f * mm * nn * m32::sqrt(d);
                    const float s_ = signf(q + hh) * m32::cbrt32(absf(q + hh));
                    const float u = signf(q - hh) * m32::cbrt32(absf(q - hh));
                    const float rx = -s_ - u - 4.f * c + 2.f * m2;
                    const float ry = sq3 * (s_ - u);
                    const float rm = hypot32(rx, ry);
                    co = (ry / m32::sqrt(rm - rx) + 2.f * g / rm - mm) / 2.f;
                }
                const float rx2 = a * co, ry2 = b * m32::sqrt(1.f - co * co);
                m.top[j] = norm2(rx2 - x, ry2 - y) * signf(y - ry2);
            }
        } break;
        case GSDF_OP_BEZIERQ2D: if constexpr (EXT) {  // :581-659  c1=(Ax,Ay,ax,ay) c2=(bx,by,cx,cy) c3=(kk,kx,kx2,a2) w2=thick/2
            const float sq3 = (float)1.7320508075688772935274463415058723669428052538103806280558069794;
            const float4 c1 = ldf4(prog, pc + 1), c2 = ldf4(prog, pc + 2), c3 = ldf4(prog, pc + 3);
            const float third = (float)(1. / 3);
            m.pushD();
#pragma unroll 1
            for (int j = 0; j < P; j++) {
                const float dx = c1.x - m.px[j], dy = c1.y - m.py[j];
                const float ky = c3.x * (2.f * c3.w + (dx * c2.x + dy * c2.y)) / 3.f;
                const float kz = c3.x * (dx * c1.z + dy * c1.w);
                const float g = ky - c3.z;
                const float q = c3.y * (2.f * c3.z - 3.f * ky) + kz;
                const float g3 = g * g * g;
                const float q2 = q * q;
                float hh = q2 + 4.f * g3;
                float res;
                if (hh >= 0.f) {
                    hh = m32::sqrt(hh);
                    float xx = 0.5f * (hh + -q), xy = 0.5f * (-hh + -q);
                    if (absf(g) < 0.001f) {
                        const float k = (1.0f - g3 / q2) * g3 / q;
                        xx = k; xy = -k - q;
                    }
                    const float ux = signf(xx) * m32::pow_frac(absf(xx), third);
                    const float uy = signf(xy) * m32::pow_frac(absf(xy), third);
                    float t = ux + uy;
                    t -= (t * (t * t + 3.0f * g) + q) / (3.0f * t * t + 3.0f * g);
                    t = clampf(t - c3.y, 0.f, 1.f);
                    const float wx = dx + t * (c2.z + t * c2.x), wy = dy + t * (c2.w + t * c2.y);
                    res = wx * wx + wy * wy;
                } else {
                    const float z = m32::sqrt(-g);
                    const float xm = m32::sqrt(0.5f + 0.5f * (q / (2.f * g * z)));  // cos_acos_3, gsdf.go:186-189
                    const float mm = xm * (xm * (xm * (xm * -0.008972f + 0.039071f) - 0.107074f) + 0.576975f) + 0.5f;
                    float nn = m32::sqrt(1.f - mm * mm);
                    nn *= sq3;
                    const float tx = clampf((mm + mm) * z - c3.y, 0.f, 1.f);
                    const float ty = clampf((-nn - mm) * z - c3.y, 0.f, 1.f);
                    const float qxx = dx + tx * (c2.z + tx * c2.x), qxy = dy + tx * (c2.w + tx * c2.y);
                    const float qyx = dx + ty * (c2.z + ty * c2.x), qyy = dy + ty * (c2.w + ty * c2.y);
                    const float ddx = qxx * qxx + qxy * qxy, ddy = qyx * qyx + qyy * qyy;
                    res = ddx < ddy ? ddx : ddy;
                }
                m.top[j] = m32::sqrt(res) - f2;
            }
        } break;
        case GSDF_OP_POLY2D: {  // :793-818; aux records (v1x,v1y,ex,ey | norm2e,v2y,_,_)
            const float4 *rec = aux + (h.y >> 2);
            const int nv = (int)h.z;
            float d[P];
            uint32_t neg = 0u;  // bit j set <=> s == -1 for point j
            {
                const float4 r0 = rec[0];
#pragma unroll
                for (int j = 0; j < P; j++) {
                    float ax = m.px[j] - r0.x, ay = m.py[j] - r0.y;
                    d[j] = ax * ax + ay * ay;
                }
            }
            for (int iv = 0; iv < nv; iv++) {
                const float4 ra = rec[2 * iv], rb = rec[2 * iv + 1];
#pragma unroll
                for (int j = 0; j < P; j++) {
                    float wx = m.px[j] - ra.x, wy = m.py[j] - ra.y;
                    float c = clampf((wx * ra.z + wy * ra.w) / rb.x, 0.f, 1.f);
                    float bx = wx - c * ra.z, by = wy - c * ra.w;
                    d[j] = minf(d[j], bx * bx + by * by);
                    bool b1 = m.py[j] >= ra.y, b2 = m.py[j] < rb.y, b3 = ra.z * wy > ra.w * wx;
                    if ((b1 && b2 && b3) || (!b1 && !b2 && !b3)) neg ^= (1u << j);
                }
            }
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float s = (neg >> j) & 1u ? -1.f : 1.f;
                m.top[j] = s * m32::sqrt(d[j]);
            }
        } break;
        // ------------------------------------------------------------------ combiners
        case GSDF_OP_MIN: {
            if (m.skip) { m.skip = false; break; }
            float a[P]; m.popBelow(a);
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = minf(a[j], m.top[j]);
        } break;
        case GSDF_OP_MAX: {
            float a[P]; m.popBelow(a);
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = maxf(a[j], m.top[j]);
        } break;
        case GSDF_OP_DIFF: {
            if (m.skip) { m.skip = false; break; }
            float a[P]; m.popBelow(a);
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = maxf(a[j], -m.top[j]);
        } break;
        case GSDF_OP_XOR: {
            float a[P]; m.popBelow(a);
#pragma unroll
            for (int j = 0; j < P; j++) { float b = m.top[j]; m.top[j] = maxf(minf(a[j], b), -maxf(a[j], b)); }
        } break;
        case GSDF_OP_SMOOTH_UNION: {  // :229-234
            if (m.skip) { m.skip = false; break; }
            float a[P]; m.popBelow(a);
#pragma unroll
            for (int j = 0; j < P; j++) {
                float b = m.top[j];
                float hh = clampf(0.5f + 0.5f * (b - a[j]) / f2, 0.f, 1.f);
                m.top[j] = (b * (1.f - hh) + a[j] * hh) - f2 * hh * (1.f - hh);
            }
        } break;
        case GSDF_OP_SMOOTH_DIFF: {  // :254-259
            float a[P]; m.popBelow(a);
#pragma unroll
            for (int j = 0; j < P; j++) {
                float b = m.top[j];
                float hh = clampf(0.5f - 0.5f * (b + a[j]) / f2, 0.f, 1.f);
                m.top[j] = (a[j] * (1.f - hh) + (-b) * hh) + f2 * hh * (1.f - hh);
            }
        } break;
        case GSDF_OP_SMOOTH_INTERSECT: {  // :279-284
            float a[P]; m.popBelow(a);
#pragma unroll
            for (int j = 0; j < P; j++) {
                float b = m.top[j];
                float hh = clampf(0.5f - 0.5f * (b - a[j]) / f2, 0.f, 1.f);
                m.top[j] = (b * (1.f - hh) + a[j] * hh) + f2 * hh * (1.f - hh);
            }
        } break;
        // ------------------------------------------------------------------ unary distance ops
        case GSDF_OP_OFFSET:
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = m.top[j] + f2;
            break;
        case GSDF_OP_ANNULUS:
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = absf(m.top[j]) - f2;
            break;
        case GSDF_OP_MIN_CONST:  // cpu_evaluators.go:364,932,1172: the folds start from a constant
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = minf(f2, m.top[j]);
            break;
        case GSDF_OP_MULDIST:
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = m.top[j] * f2;
            break;
        case GSDF_OP_SHELL_EXIT:
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = f2 * (absf(m.top[j]) - f2);
            break;
        case GSDF_OP_ADD_BELOW: {
            float a[P]; m.popBelow(a);
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = m.top[j] + a[j];
        } break;
        case GSDF_OP_EXTRUDE_EXIT: {  // :524-529
            float a[P]; m.popBelow(a);
#pragma unroll
            for (int j = 0; j < P; j++) {
                float d = m.top[j], wy = a[j];
                m.top[j] = minf(0.f, maxf(d, wy)) + hypot32(maxf(d, 0.f), maxf(wy, 0.f));
            }
        } break;
        case GSDF_OP_MAX_BELOW: {  // threads.go:176-180
            float a[P]; m.popBelow(a);
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = maxf(m.top[j], a[j]);
        } break;
        // ------------------------------------------------------------------ position stack
        case GSDF_OP_PUSH_POS: m.pushPos(m.px, m.py, m.pz); break;
        case GSDF_OP_POP_POS: m.psp--; m.loadPos(m.psp); break;
        case GSDF_OP_PEEK_POS: m.loadPos(m.psp - 1); break;
        // ------------------------------------------------------------------ position transforms
        case GSDF_OP_TRANSLATE: {  // :470, :980
            const float4 c = ldf4(prog, pc + 1);
#pragma unroll
            for (int j = 0; j < P; j++) { m.px[j] = m.px[j] - c.x; m.py[j] = m.py[j] - c.y; m.pz[j] = m.pz[j] - c.z; }
        } break;
        case GSDF_OP_SCALE_POS:  // :300-302
#pragma unroll
            for (int j = 0; j < P; j++) { m.px[j] = f2 * m.px[j]; m.py[j] = f2 * m.py[j]; m.pz[j] = f2 * m.pz[j]; }
            break;
        case GSDF_OP_SYMMETRY:  // :314
#pragma unroll
            for (int j = 0; j < P; j++) {
                if (h.y & 1u) m.px[j] = absf(m.px[j]);
                if (h.y & 2u) m.py[j] = absf(m.py[j]);
                if (h.y & 4u) m.pz[j] = absf(m.pz[j]);
            }
            break;
        case GSDF_OP_TRANSFORM: {  // :488-498 MulPosition, w=1
            const float4 r0 = ldf4(prog, pc + 1), r1 = ldf4(prog, pc + 2), r2 = ldf4(prog, pc + 3);
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = m.px[j], y = m.py[j], z = m.pz[j];
                m.px[j] = r0.x * x + r0.y * y + r0.z * z + r0.w;
                m.py[j] = r1.x * x + r1.y * y + r1.z * z + r1.w;
                m.pz[j] = r2.x * x + r2.y * y + r2.z * z + r2.w;
            }
        } break;
        case GSDF_OP_ROTATE2D: {  // :1186
            const float4 c = ldf4(prog, pc + 1);
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = m.px[j], y = m.py[j];
                m.px[j] = c.x * x + c.y * y;
                m.py[j] = c.z * x + c.w * y;
            }
        } break;
        case GSDF_OP_TWIST:  // :1257
#pragma unroll
            for (int j = 0; j < P; j++) {
                float s, c;
                m32::sincos(f2 * m.pz[j], s, c);
                float x = m.px[j], y = m.py[j];
                m.px[j] = c * x - s * y;
                m.py[j] = s * x + c * y;
            }
            break;
        case GSDF_OP_ELONGATE: {  // :399-417
            const float4 c = ldf4(prog, pc + 1);
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float qx = absf(m.px[j]) - c.x, qy = absf(m.py[j]) - c.y, qz = absf(m.pz[j]) - c.z;
                m.top[j] = minf(maxf(qx, maxf(qy, qz)), 0.f);
                m.px[j] = maxf(qx, 0.f); m.py[j] = maxf(qy, 0.f); m.pz[j] = maxf(qz, 0.f);
            }
        } break;
        case GSDF_OP_ELONGATE2D:  // :1228-1246
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) {
                float qx = absf(m.px[j]) - f2, qy = absf(m.py[j]) - f3;
                m.top[j] = minf(maxf(qx, qy), 0.f);
                m.px[j] = maxf(qx, 0.f); m.py[j] = maxf(qy, 0.f);
            }
            break;
        case GSDF_OP_ARRAY_VAR: {  // :368-383
            const float4 s = ldf4(prog, pc + 1), n = ldf4(prog, pc + 2);
            const float fi = (float)(h.y & 1u), fj = (float)((h.y >> 1) & 1u), fk = (float)((h.y >> 2) & 1u);
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = m.px[j], y = m.py[j], z = m.pz[j];
                float idx = roundf(x / s.x), idy = roundf(y / s.y), idz = roundf(z / s.z);
                float ox = signf(x - s.x * idx), oy = signf(y - s.y * idy), oz = signf(z - s.z * idz);
                float rx = clampf(idx + fi * ox, 0.f, n.x), ry = clampf(idy + fj * oy, 0.f, n.y), rz = clampf(idz + fk * oz, 0.f, n.z);
                m.px[j] = x - s.x * rx; m.py[j] = y - s.y * ry; m.pz[j] = z - s.z * rz;
            }
        } break;
        case GSDF_OP_ARRAY2D_VAR: {  // :936-949  c=(sx,sy,nx-1,ny-1)
            const float4 c = ldf4(prog, pc + 1);
            const float fi = (float)(h.y & 1u), fj = (float)((h.y >> 1) & 1u);
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = m.px[j], y = m.py[j];
                float idx = roundf(x / c.x), idy = roundf(y / c.y);
                float ox = signf(x - c.x * idx), oy = signf(y - c.y * idy);
                float rx = clampf(idx + fi * ox, 0.f, c.z), ry = clampf(idy + fj * oy, 0.f, c.w);
                m.px[j] = x - c.x * rx; m.py[j] = y - c.y * ry;
            }
        } break;
        case GSDF_OP_CIRC_ENTER: {  // :1056-1078  c=(angle,ncirc,ninsm1,table)
            const float4 c = ldf4(prog, pc + 1);
            const uint32_t tab = __float_as_uint(c.w);
            float x0[P], y0[P];
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = m.px[j], y = m.py[j];
                float pangle = m32::atan2(y, x);
                float id = floorf(pangle / c.x);
                if (id < 0.f) id += c.y;
                float i0, i1;
                if (id >= c.z) { i0 = c.z; i1 = 0.f; } else { i0 = id; i1 = id + 1.f; }
                float s0, c0, s1, c1;
                // Sincos(angle * i) for i = 0..ncirc from the table the library appended to the side buffer (capi.cu
                // augment_program; word 3 of the operands = its position + 1, 0 = no table); anything else (NaN) is computed
                if (tab && i0 >= 0.f && i0 <= c.y && i1 >= 0.f && i1 <= c.y) {
                    const float2 *T = reinterpret_cast<const float2 *>(aux + (tab - 1u));
                    const float2 t0 = T[(int)i0], t1 = T[(int)i1];
                    s0 = t0.x; c0 = t0.y; s1 = t1.x; c1 = t1.y;
                } else {
                    m32::sincos(c.x * i0, s0, c0);
                    m32::sincos(c.x * i1, s1, c1);
                }
                x0[j] = c0 * x + s0 * y; y0[j] = -s0 * x + c0 * y;
                m.px[j] = c1 * x + s1 * y; m.py[j] = -s1 * x + c1 * y;
            }
            m.pushPos(x0, y0, m.pz);
        } break;
        case GSDF_OP_CULL_UB2D: {  // box guards (gsdf_program.h): upper bound of the union from anchor points on the outlines
            const float4 *an = aux + (h.y >> 2);
            const int npairs = (int)(h.z >> 1);
            float best[P];
#pragma unroll
            for (int j = 0; j < P; j++) best[j] = 3.0e38f;
            for (int q = 0; q < npairs; q++) {
                const float4 v = an[q];
#pragma unroll
                for (int j = 0; j < P; j++) {
                    const float ax = m.px[j] - v.x, ay = m.py[j] - v.y, bx = m.px[j] - v.z, by = m.py[j] - v.w;
                    best[j] = minf(best[j], minf(ax * ax + ay * ay, bx * bx + by * by));
                }
            }
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = m32::sqrt(best[j]) * 1.0001f + f3;
        } break;
        case GSDF_OP_BBOX_GUARD2D: {
            const float4 bx = ldf4(prog, pc + 1);
            bool dead = true;
#pragma unroll
            for (int j = 0; j < P; j++) {
                const float dx = maxf(maxf(bx.x - m.px[j], m.px[j] - bx.z), 0.f), dy = maxf(maxf(bx.y - m.py[j], m.py[j] - bx.w), 0.f);
                const float w = m32::sqrt(dx * dx + dy * dy) * 0.9999f - f3;
                // the box bounds the operand from below only OUTSIDE the box: a point inside it never votes for the skip
                dead &= (dx > 0.f || dy > 0.f) && guard_dead(h.y & 0xffu, w, m.top[j], 0.f);
            }
            if (GSDF_GUARD_VOTE(dead)) { m.skip = true; pc = (int)(h.y >> 8); return true; }
        } break;
        case GSDF_OP_EXTRUDE_ENTER: {  // :524-527  f2=h/2
            if (h.y & 0xffu) {
                bool dead = true;
#pragma unroll
                for (int j = 0; j < P; j++) dead &= guard_dead(h.y & 0xffu, absf(m.pz[j]) - f2, m.top[j], f3);
                if (GSDF_GUARD_VOTE(dead)) { m.skip = true; pc = (int)(h.y >> 8); return true; }
            }
            m.pushD();
#pragma unroll
            for (int j = 0; j < P; j++) m.top[j] = absf(m.pz[j]) - f2;
        } break;
        case GSDF_OP_REVOLVE:  // :545-547
#pragma unroll
            for (int j = 0; j < P; j++) { m.px[j] = hypot32(m.px[j], m.pz[j]) - f2; }
            break;
        case GSDF_OP_SCREW_ENTER: {  // threads.go:156-170,198-202  c=(pitch,lead,L/2,tanTaper)
            const float4 c = ldf4(prog, pc + 1);
            if (h.y & 0xffu) {
                bool dead = true;
#pragma unroll
                for (int j = 0; j < P; j++) dead &= guard_dead(h.y & 0xffu, absf(m.pz[j]) - c.z, m.top[j], f3);
                if (GSDF_GUARD_VOTE(dead)) { m.skip = true; pc = (int)(h.y >> 8); return true; }
            }
            m.pushD();
#ifdef GSDF_RXY
            float r[P];
            m.radius(h.z, r);  // SCREW_ENTER keeps its radius flags in word 2 (word 1 is the slab guard)
#endif
#pragma unroll
            for (int j = 0; j < P; j++) {
                float x = m.px[j], y = m.py[j], z = m.pz[j];
#ifdef GSDF_RXY
                float yy = r[j];
#else
                float yy = hypot32(x, y);
#endif
                yy += z * c.w;
                float theta = m32::atan2(y, x);
                float zz = z + c.y * theta / m32::kTwoPiF;
                float sx = zz + c.x / 2.f;
                float t = sx / c.x;
                m.px[j] = c.x * (t - floorf(t)) - c.x / 2.f;
                m.py[j] = yy;
                m.top[j] = absf(z) - c.z;
            }
        } break;
        default:
            return false;  // unknown opcode: rejected at gsdf_program_create, never reached
        }
        pc += len;
    }
    return true;
}

// Runs the program at the P positions already loaded in m.px/py/pz; result in m.top.
template <int P, bool EXT>
__device__ __forceinline__ void run_program(Machine<P> &m, const uint4 *__restrict__ prog, const float4 *__restrict__ aux) {
    int pc = 0;
    for (;;) {
#ifdef GSDF_LOCKSTEP
        __syncthreads();  // keep the CTA's warps on the same opcode body: one instruction-cache stream per CTA
#endif
        if (!exec_one<P, EXT>(m, prog[pc], prog, pc, aux)) return;
    }
}

}  // namespace gsdfk
