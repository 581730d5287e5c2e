// Test-only stand-in for <cuda_runtime.h>: lets g++ compile gsdf_b200/csrc/interp.cuh + math32.cuh as HOST code, so that
// the very source the GPU runs can be executed on the CPU (tests/host_interp/host_interp.cpp). One "thread" per machine:
// barriers are no-ops and the CTA-wide vote of a guard is the vote of the thread's own P points.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#define __host__
#define __device__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint4 { uint32_t x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline void __syncthreads() {}
static inline int __syncthreads_and(int p) { return p; }

// fminf / fmaxf with the semantics of the device instructions (PTX min.f32 / max.f32: a NaN operand yields the other
// operand, and -0.0 orders below +0.0). The host C library leaves the sign of min(+0, -0) unspecified, and a distance
// of -0.0 vs +0.0 is a visible bit difference against the oracle (Go's math32.Min / Max order the zeros the same way).
static inline float gsdf_shim_fminf(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.f && b == 0.f) return std::signbit(a) ? a : b;
    return a < b ? a : b;
}
static inline float gsdf_shim_fmaxf(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.f && b == 0.f) return std::signbit(a) ? b : a;
    return a > b ? a : b;
}
#define fminf(a, b) gsdf_shim_fminf(a, b)
#define fmaxf(a, b) gsdf_shim_fmaxf(a, b)
