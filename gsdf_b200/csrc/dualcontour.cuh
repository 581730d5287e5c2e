// dualcontour.cuh -- sm_100a kernels of the dual-contouring renderer (glrender/dual_contour.go,
// glrender/dual_contour_vertexplacement.go, gleval.NormalsCentralDiff gleval/gleval.go:53-108).
//
// The reference keeps the surviving level-1 cubes in a slice (octree BFS order) plus a map[i3.Vec]int; here the N^3
// cube origins are evaluated in that same order (index = "BFS key": the per-level child indices of i3.Cube.Octree(),
// most significant level first), the prune flags are turned into cube indices by an exclusive scan -- which IS the
// cubeMap: index[key] -- and every later stage is one thread per surviving cube. Stages:
//   k_eval<4,GenDC mode 0>   distance at every cube origin                      (Reset: DecomposeBFS + octreePrunea)
//   k_dc_flags / scan / k_dc_compact   |d| < 2 res -> ordered cube list + index map
//   k_eval<4,GenDC mode 1>   origin + three edge ends of each kept cube         (RenderAll, dual_contour.go:85-107)
//   k_eval<4,GenDC mode 2>   central differences at the edge intersections      (PlaceVertices, :28-50)
//   k_dc_place               neighbour gather, QEF rows, float64 MGS least squares (vertexplacement.go:52-223)
//   scan / k_dc_emit         quads -> triangles in cube order                   (dual_contour.go:152-218)
// Every float32/float64 operation is individually rounded (-fmad=false) and ordered as in the Go code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dc_gen.cuh"

namespace gsdfk {

__global__ void __launch_bounds__(256) k_dc_flags(const float *__restrict__ dist, uint32_t n, float res, uint32_t *__restrict__ flags) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) flags[i] = dc_kept(dist[i], res) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_dc_compact(const float *__restrict__ dist, const uint32_t *__restrict__ eidx, uint32_t n, float res,
                                                   uint32_t *__restrict__ cubekey) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (dc_kept(dist[i], res)) cubekey[eidx[i]] = i;
}

struct DCArgs {
    DCGrid G;
    const float *dist;        // per key
    const uint32_t *eidx;     // per key: cube index where kept (the cubeMap)
    const uint32_t *cubekey;  // per cube
    uint32_t ncubes;
    const float4 *dc4;
    const float *nrm;         // 9 floats per cube (3 edges x 3 components), LSQ placers only
    float3 *fin;              // FinalVertex per cube
    uint32_t *qcount;         // quads per cube -> exclusive offsets after the scan
    int placer;               // GSDF_DC_*
    float sqrtLambda;
    float *tris;
    unsigned long long *with_neighbors;
    uint32_t key0, key1;      // cubes with key in [key0, key1) are OWNED by this part: only they emit quads / are counted
};

// cubeMap[iv] (dual_contour.go:98): cube index of cell (i,j,k) or -1
__device__ __forceinline__ int dc_lookup(const DCArgs &A, int i, int j, int k) {
    const int N = 1 << A.G.bits;
    if ((unsigned)i >= (unsigned)N || (unsigned)j >= (unsigned)N || (unsigned)k >= (unsigned)N) return -1;
    const uint32_t key = dc_key(i, j, k, A.G.bits);
    return dc_kept(A.dist[key], A.G.res) ? (int)A.eidx[key] : -1;
}
// EdgeNeighborsX/Y/Z (dual_contour.go:282-298) in cell units: [axis][q][xyz]
__device__ __constant__ int8_t kDcEnb[3][4][3] = {
    {{0, -1, -1}, {0, 0, -1}, {0, 0, 0}, {0, -1, 0}},
    {{-1, 0, -1}, {-1, 0, 0}, {0, 0, 0}, {0, 0, -1}},
    {{-1, -1, 0}, {0, -1, 0}, {0, 0, 0}, {-1, 0, 0}},
};

// leastSquaresMGS64 (dual_contour_vertexplacement.go:148-223): K x 3 system in float64, modified Gram-Schmidt.
__device__ void dc_lsq_mgs64(int K, double (*Q)[3], const double *b64, float (&out)[3]) {
    out[0] = out[1] = out[2] = 0.f;
    if (K < 3) return;
    double R[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int j = 0; j < 3; j++) {
        for (int i = 0; i < j; i++) {
            double dot = 0;
            for (int k = 0; k < K; k++) dot += Q[k][i] * Q[k][j];
            R[i][j] = dot;
            for (int k = 0; k < K; k++) Q[k][j] -= dot * Q[k][i];
        }
        double normSq = 0;
        for (int k = 0; k < K; k++) normSq += Q[k][j] * Q[k][j];
        const double norm = sqrt(normSq);
        R[j][j] = norm;
        if (norm > 1e-14) {
            const double inv = 1.0 / norm;
            for (int k = 0; k < K; k++) Q[k][j] *= inv;
        }
    }
    double Qtb[3] = {0, 0, 0};
    for (int j = 0; j < 3; j++)
        for (int k = 0; k < K; k++) Qtb[j] += Q[k][j] * b64[k];
    double x[3];
    for (int i = 2; i >= 0; i--) {
        x[i] = Qtb[i];
        for (int k = i + 1; k < 3; k++) x[i] -= R[i][k] * x[k];
        if (R[i][i] > 1e-14) x[i] /= R[i][i]; else x[i] = 0;
    }
    out[0] = (float)x[0]; out[1] = (float)x[1]; out[2] = (float)x[2];
}

// One thread per cube: Neighbors (dual_contour.go:118-146, entries ordered by (cube index, axis) as the reference's
// append order produces), PlaceVertices, and the number of quads the cube will emit.
__global__ void __launch_bounds__(128) k_dc_place(DCArgs A) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= A.ncubes) return;
    int ci, cj, ck;
    dc_unkey(A.cubekey[c], A.G.bits, ci, cj, ck);
    const float4 own = A.dc4[c];
    const float res = A.G.res;
    const float3 co = dc_origin(A.G, ci, cj, ck);
    const bool act[3] = {dc_active(own.x, own.y), dc_active(own.x, own.z), dc_active(own.x, own.w)};
    // quads this cube emits: an active edge whose four surrounding cubes all exist (dual_contour.go:158-171)
    uint32_t nq = 0;
    for (int a = 0; a < 3; a++) {
        if (!act[a]) continue;
        bool all = true;
        for (int q = 0; q < 4 && all; q++) all = dc_lookup(A, ci + kDcEnb[a][q][0], cj + kDcEnb[a][q][1], ck + kDcEnb[a][q][2]) >= 0;
        nq += all ? 1u : 0u;
    }
    const uint32_t ckey = A.cubekey[c];
    const bool owned = ckey >= A.key0 && ckey < A.key1;
    A.qcount[c] = owned ? nq : 0u;
    // entries (n, a): cube n whose active a-edge touches this voxel, i.e. this cube = n + ENB[a][q]
    uint32_t ent[12];
    int nnb = 0;
    for (int a = 0; a < 3; a++)
        for (int q = 0; q < 4; q++) {
            const int n = dc_lookup(A, ci - kDcEnb[a][q][0], cj - kDcEnb[a][q][1], ck - kDcEnb[a][q][2]);
            if (n < 0) continue;
            const float4 d = A.dc4[n];
            const float ed = a == 0 ? d.y : (a == 1 ? d.z : d.w);
            if (dc_active(d.x, ed)) ent[nnb++] = ((uint32_t)n << 2) | (uint32_t)a;
        }
    for (int i = 1; i < nnb; i++) {  // insertion sort by (n, a)
        const uint32_t v = ent[i];
        int j = i - 1;
        while (j >= 0 && ent[j] > v) { ent[j + 1] = ent[j]; j--; }
        ent[j + 1] = v;
    }
    if (nnb == 0) { A.fin[c] = co; return; }  // default FinalVertex = cube origin (dual_contour.go:114)
    if (owned) atomicAdd(A.with_neighbors, 1ull);
    // contribution of entry (n, a): the edge's linear zero crossing
    auto contrib = [&](uint32_t n, int a) {
        int i, j, k;
        dc_unkey(A.cubekey[n], A.G.bits, i, j, k);
        const float3 o = dc_origin(A.G, i, j, k);
        const float4 d = A.dc4[n];
        const float s = res * dc_isect(d.x, a == 0 ? d.y : (a == 1 ? d.z : d.w));
        return make_float3(o.x + (a == 0 ? s : 0.f), o.y + (a == 1 ? s : 0.f), o.z + (a == 2 ? s : 0.f));
    };
    if (A.placer == 0) {  // DualContourNaive (dual_contour_test.go:358-389)
        float3 sum = make_float3(0.f, 0.f, 0.f);
        for (int r = 0; r < nnb; r++) {
            const float3 p = contrib(ent[r] >> 2, (int)(ent[r] & 3u));
            sum.x += p.x; sum.y += p.y; sum.z += p.z;
        }
        const float inv = 1.0f / (float)nnb;
        A.fin[c] = make_float3(sum.x * inv, sum.y * inv, sum.z * inv);
        return;
    }
    // DualContourLeastSquares: rows = own active edges, then every Neighbors entry (own edges appear again there)
    double Q[18][3], b64[18];
    int K = 0;
    float3 mean = make_float3(0.f, 0.f, 0.f);
    const float invRes = 1.0f / res;
    auto row = [&](float3 p, const float *n) {
        const float qx = invRes * (p.x - co.x), qy = invRes * (p.y - co.y), qz = invRes * (p.z - co.z);
        const float bb = n[0] * qx + n[1] * qy + n[2] * qz;
        Q[K][0] = (double)n[0]; Q[K][1] = (double)n[1]; Q[K][2] = (double)n[2];
        b64[K] = (double)bb;
        K++;
        mean.x += p.x; mean.y += p.y; mean.z += p.z;
    };
    for (int a = 0; a < 3; a++)
        if (act[a]) row(contrib(c, a), A.nrm + 9 * (size_t)c + 3 * a);
    for (int r = 0; r < nnb; r++) {
        const uint32_t n = ent[r] >> 2;
        const int a = (int)(ent[r] & 3u);
        row(contrib(n, a), A.nrm + 9 * (size_t)n + 3 * a);
    }
    const float invn = 1.f / (float)K;  // vertMean (vertexplacement.go:140-145)
    mean = make_float3(invn * mean.x, invn * mean.y, invn * mean.z);
    const float bx = invRes * (mean.x - co.x), by = invRes * (mean.y - co.y), bz = invRes * (mean.z - co.z);
    const float sl = A.sqrtLambda;
    Q[K][0] = (double)sl; Q[K][1] = 0; Q[K][2] = 0; b64[K++] = (double)(sl * bx);
    Q[K][0] = 0; Q[K][1] = (double)sl; Q[K][2] = 0; b64[K++] = (double)(sl * by);
    Q[K][0] = 0; Q[K][1] = 0; Q[K][2] = (double)sl; b64[K++] = (double)(sl * bz);
    float x[3];
    dc_lsq_mgs64(K, Q, b64, x);
#pragma unroll
    for (int a = 0; a < 3; a++) x[a] = x[a] < -0.1f ? -0.1f : (x[a] > 1.1f ? 1.1f : x[a]);  // ClampElem, 10 % relaxation
    A.fin[c] = make_float3(res * x[0] + co.x, res * x[1] + co.y, res * x[2] + co.z);
}

// One thread per cube: its quads as two triangles each, at the scanned offset (dual_contour.go:152-218).
__global__ void __launch_bounds__(128) k_dc_emit(DCArgs A) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= A.ncubes) return;
    int ci, cj, ck;
    dc_unkey(A.cubekey[c], A.G.bits, ci, cj, ck);
    const float4 own = A.dc4[c];
    uint64_t o = (uint64_t)A.qcount[c];
    if (A.cubekey[c] < A.key0 || A.cubekey[c] >= A.key1) return;
    for (int a = 0; a < 3; a++) {
        const float ed = a == 0 ? own.y : (a == 1 ? own.z : own.w);
        if (!dc_active(own.x, ed)) continue;
        float3 quad[4];
        bool all = true;
        for (int q = 0; q < 4 && all; q++) {
            const int n = dc_lookup(A, ci + kDcEnb[a][q][0], cj + kDcEnb[a][q][1], ck + kDcEnb[a][q][2]);
            if (n < 0) all = false; else quad[q] = A.fin[n];
        }
        if (!all) continue;
        if (ed - own.x < 0.f) {  // FlipX/Y/Z (dual_contour.go:278-280)
            const float3 t0 = quad[0], t1 = quad[1];
            quad[0] = quad[3]; quad[1] = quad[2]; quad[2] = t1; quad[3] = t0;
        }
        float *dst = A.tris + 18 * o;
        const float3 v[6] = {quad[0], quad[1], quad[2], quad[2], quad[3], quad[0]};
#pragma unroll
        for (int t = 0; t < 6; t++) { dst[3 * t] = v[t].x; dst[3 * t + 1] = v[t].y; dst[3 * t + 2] = v[t].z; }
        o++;
    }
}

}  // namespace gsdfk
