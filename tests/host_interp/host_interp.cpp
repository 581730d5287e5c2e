// Test infrastructure: the device interpreter's own source (gsdf_b200/csrc/interp.cuh, math32.cuh) compiled for the host.
// g++ -O2 -ffp-contract=off -I tests/host_interp/shim -I gsdf_b200/csrc ... -> every float32 operation rounds individually,
// like nvcc -fmad=false with the _rn intrinsics on the device. What differs from the GPU: the math32 helpers take their
// host branches (sqrtf, a / b instead of __fsqrt_rn, __fdiv_rn -- the same IEEE operations), one machine at a time
// instead of 512 threads, and a guard's vote spans the 4 points of one machine instead of a 2048-point tile (a finer
// tiling; guards must be value-preserving under ANY tiling).
#include <cuda_runtime.h>  // the shim
#include <vector>
#include <algorithm>

#include "interp.cuh"

extern "C" int host_interp_eval(const uint32_t *chunks, const float *aux, uint32_t dslots, uint32_t pslots, int dim, const float *pos,
                                float *out, size_t n, int ext) {
    using namespace gsdfk;
    constexpr int P = 4;
    if (n == 0) return 0;
    std::vector<float> dstk((size_t)P * (dslots + 1)), pstk((size_t)P * 3 * (pslots + 1));
#ifdef GSDF_RXY
    std::vector<float> rxy(P);  // the experimental radius cache (gsdf_program.h, "Radius reuse"): one float per point
#endif
    for (size_t i = 0; i < n; i += P) {
        Machine<P> m;
        m.init(dstk.data(), pstk.data(), 1);
#ifdef GSDF_RXY
        m.rxy = rxy.data();
#endif
        for (int j = 0; j < P; j++) {
            const size_t idx = std::min(i + j, n - 1);
            m.px[j] = pos[dim * idx];
            m.py[j] = pos[dim * idx + 1];
            m.pz[j] = dim == 3 ? pos[dim * idx + 2] : 0.f;
        }
        if (ext) run_program<P, true>(m, reinterpret_cast<const uint4 *>(chunks), reinterpret_cast<const float4 *>(aux));
        else run_program<P, false>(m, reinterpret_cast<const uint4 *>(chunks), reinterpret_cast<const float4 *>(aux));
        for (int j = 0; j < P && i + j < n; j++) out[i + j] = m.top[j];
    }
    return 0;
}

// math32.cuh's elementary functions (host branches) over arrays, for dense comparisons with the oracle's restatement.
extern "C" void host_m32_unary(int fn, const float *x, float *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        float s, c;
        switch (fn) {
        case 0: out[i] = m32::sin(x[i]); break;
        case 1: out[i] = m32::cos(x[i]); break;
        case 2: out[i] = m32::tan(x[i]); break;
        case 3: out[i] = m32::atan(x[i]); break;
        case 4: out[i] = m32::acos(x[i]); break;
        case 5: out[i] = m32::cbrt32(x[i]); break;
        case 6: out[i] = m32::log32(x[i]); break;
        case 7: out[i] = m32::exp32(x[i]); break;
        case 8: m32::sincos(x[i], s, c); out[i] = s; break;
        case 9: m32::sincos(x[i], s, c); out[i] = c; break;
        case 10: out[i] = m32::sqrt(x[i]); break;
        default: out[i] = 0.f;
        }
    }
}
extern "C" void host_m32_binary(int fn, const float *x, const float *y, float *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        switch (fn) {
        case 0: out[i] = m32::hypot32(x[i], y[i]); break;
        case 1: out[i] = m32::atan2(x[i], y[i]); break;
        case 2: out[i] = m32::minf(x[i], y[i]); break;
        case 3: out[i] = m32::maxf(x[i], y[i]); break;
        default: out[i] = 0.f;
        }
    }
}
