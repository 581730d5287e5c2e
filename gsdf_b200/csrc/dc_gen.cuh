// dc_gen.cuh -- dual-contouring grid keys and the generator that feeds the interpreter kernel (shared by eval.cu, which
// instantiates k_eval<4,GenDC>, and dualcontour.cu). See dualcontour.cuh for the stage overview.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsdfk {

struct DCGrid {
    float ox, oy, oz, res;
    int bits;        // levels - 1: N = 1 << bits cubes per axis
    uint32_t ncell;  // N^3
};

// i3.Cube.Octree() child order = Bourke corner order (0,0,0)(1,0,0)(1,1,0)(0,1,0)(0,0,1)(1,0,1)(1,1,1)(0,1,1)
__host__ __device__ __forceinline__ void dc_unkey(uint32_t key, int bits, int &i, int &j, int &k) {
    int x = 0, y = 0, z = 0;
    for (int b = bits - 1; b >= 0; b--) {
        const int c = (int)((key >> (3 * b)) & 7u);
        const int zb = c >> 2, r = c & 3, yb = r >> 1, xb = yb ? 3 - r : r;
        x |= xb << b; y |= yb << b; z |= zb << b;
    }
    i = x; j = y; k = z;
}
__host__ __device__ __forceinline__ uint32_t dc_key(int i, int j, int k, int bits) {
    uint32_t key = 0;
    for (int b = bits - 1; b >= 0; b--) {
        const int xb = (i >> b) & 1, yb = (j >> b) & 1, zb = (k >> b) & 1;
        key = (key << 3) | (uint32_t)(zb * 4 + (yb ? 3 - xb : xb));
    }
    return key;
}
__device__ __forceinline__ float3 dc_origin(const DCGrid &G, int i, int j, int k) {  // ms3.Octree.CubeOrigin at level 1
    return make_float3(G.ox + G.res * (float)i, G.oy + G.res * (float)j, G.oz + G.res * (float)k);
}
__device__ __forceinline__ bool dc_kept(float d, float res) { return !(fabsf(d) >= res * 2.f); }  // octreerenderer.go:271-274, mult 2
__device__ __forceinline__ bool dc_active(float o, float e) { return (__float_as_uint(o) >> 31) != (__float_as_uint(e) >> 31); }  // dual_contour.go:266-274
__device__ __forceinline__ float dc_isect(float o, float e) { return -o / (e - o); }                                              // :275-277

// Generator for the three interpreter passes (one k_eval instantiation; `mode` is launch-uniform).
struct GenDC {
    static constexpr bool kTileSkip = true;
    int mode;
    DCGrid G;
    int blo[3], bhi[3];        // mode 0: cube origins outside [blo, bhi) are not needed by this part (multi-GPU octant split)
    int clip;                  // 0: the box is the whole grid (single part), no test needed
    __device__ bool dead(uint64_t w) const {
        if (mode != 0 || !clip) return false;
        bool out = true;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const uint32_t key = min((uint32_t)(4 * w + t), G.ncell - 1);
            int i, j, k;
            dc_unkey(key, G.bits, i, j, k);
            out &= (i < blo[0] || i >= bhi[0] || j < blo[1] || j >= bhi[1] || k < blo[2] || k >= bhi[2]);
        }
        return out;
    }
    __device__ void store_dead(uint64_t w) const {  // +inf = pruned (|d| >= 2 res)
        const float inf = __int_as_float(0x7f800000);
        if (4 * w + 4 <= G.ncell) reinterpret_cast<float4 *>(dist)[w] = make_float4(inf, inf, inf, inf);
        else
            for (int t = 0; t < 4; t++) if (4 * w + t < G.ncell) dist[4 * w + t] = inf;
    }
    float *dist;               // mode 0 out: dist[key]
    const uint32_t *cubekey;   // modes 1,2: key of cube e
    uint32_t ncubes;
    float4 *dc4;               // mode 1 out / mode 2 in: {OrigDist, XDist, YDist, ZDist}
    float step;                // mode 2: NormalsCentralDiff step (already halved)
    float *nrm;                // mode 2 out: nrm[3*(3e+a) + dim]
    __device__ uint64_t work_items() const {
        if (mode == 0) return ((uint64_t)G.ncell + 3) / 4;
        if (mode == 1) return ncubes;
        return (uint64_t)ncubes * 6;  // (e, axis, half)
    }
    __device__ void load(uint64_t w, float (&x)[4], float (&y)[4], float (&z)[4]) const {
        if (mode == 0) {
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint32_t key = min((uint32_t)(4 * w + t), G.ncell - 1);
                int i, j, k;
                dc_unkey(key, G.bits, i, j, k);
                const float3 p = dc_origin(G, i, j, k);
                x[t] = p.x; y[t] = p.y; z[t] = p.z;
            }
            return;
        }
        const uint32_t e = mode == 1 ? (uint32_t)w : (uint32_t)(w / 6);
        int i, j, k;
        dc_unkey(cubekey[e], G.bits, i, j, k);
        const float3 o = dc_origin(G, i, j, k);
        if (mode == 1) {  // dual_contour.go:90-96
            x[0] = o.x; y[0] = o.y; z[0] = o.z;
            x[1] = o.x + G.res; y[1] = o.y + 0.f; z[1] = o.z + 0.f;
            x[2] = o.x + 0.f; y[2] = o.y + G.res; z[2] = o.z + 0.f;
            x[3] = o.x + 0.f; y[3] = o.y + 0.f; z[3] = o.z + G.res;
            return;
        }
        const int a = (int)((w % 6) >> 1), half = (int)(w & 1);
        const float4 d = dc4[e];
        const float ed = a == 0 ? d.y : (a == 1 ? d.z : d.w);
        float3 p = o;  // inactive edges are never read back: evaluate them at the (finite) cube origin
        if (dc_active(d.x, ed)) {
            const float s = G.res * dc_isect(d.x, ed);  // vertexplacement.go:33-37
            p = make_float3(o.x + (a == 0 ? s : 0.f), o.y + (a == 1 ? s : 0.f), o.z + (a == 2 ? s : 0.f));
        }
        // gleval.go:73-90: p + h, p - h per dimension. half 0 carries dims x,y; half 1 carries z (slots 2,3 repeat it).
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int dim = half ? 2 : (t >> 1);
            const float h = (t & 1) ? -step : step;
            x[t] = p.x + (dim == 0 ? h : (t & 1) ? -0.f : 0.f);
            y[t] = p.y + (dim == 1 ? h : (t & 1) ? -0.f : 0.f);
            z[t] = p.z + (dim == 2 ? h : (t & 1) ? -0.f : 0.f);
        }
    }
    __device__ void store(uint64_t w, const float (&d)[4]) const {
        if (mode == 0) {
            if (4 * w + 4 <= G.ncell) reinterpret_cast<float4 *>(dist)[w] = make_float4(d[0], d[1], d[2], d[3]);
            else
                for (int t = 0; t < 4; t++) if (4 * w + t < G.ncell) dist[4 * w + t] = d[t];
            return;
        }
        if (mode == 1) { dc4[w] = make_float4(d[0], d[1], d[2], d[3]); return; }
        const uint64_t ea = w >> 1;  // 3e + a
        if (w & 1) nrm[3 * ea + 2] = d[0] - d[1];
        else { nrm[3 * ea] = d[0] - d[1]; nrm[3 * ea + 1] = d[2] - d[3]; }
    }
};

}  // namespace gsdfk
