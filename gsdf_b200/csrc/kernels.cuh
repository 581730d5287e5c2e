// kernels.cuh -- sm_100a kernels of the SDF evaluate + mesh path.
//
//   k_eval<P,Gen>      one persistent grid; each thread interprets the node program at P points produced by a
//                      generator functor (AoS point lists, the dense lattice, a compacted quad list, block centres,
//                      image rows) and hands the distances to the generator's sink.
//   k_compact_quads    octree level-3 prune -> compacted list of 4-corner lattice quads that still need evaluating
//   k_mc_count/emit    marching-cubes classification per 32-cell row segment with a warp inclusive scan; triangle
//                      offsets come from an exclusive scan over segment counts so output order is the reference
//                      FlatRenderer's (cell index x fastest, flatrenderer.go:208-212) and fully deterministic.
//   k_scan_*           three-kernel exclusive scan over segment counts.
//   k_stl_pack         glrender/stl.go:33-61 record packing.
//
// The node program (+ side buffer when it fits) is staged into shared memory once per CTA by a 1-D bulk async copy
// (cp.async.bulk.shared::cluster.global, completion on an mbarrier: SASS UBLKCP / SYNCS).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "interp.cuh"
#include "mc_tables.cuh"

namespace gsdfk {

constexpr int kThreads = 256;

struct ProgView {
    const uint4 *g_prog;     // device: program chunks followed by aux (16-byte aligned)
    uint32_t prog_bytes;     // bytes of chunks
    uint32_t aux_bytes;      // bytes of aux that follow the chunks
    uint32_t stage_aux;      // 1: aux is staged to smem with the program; 0: read from global
    uint32_t dslots, pslots; // stack slots
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Stage `bytes` (multiple of 16) from global to shared with one bulk async copy; all threads return after it landed.
__device__ __forceinline__ void bulk_stage(void *s_dst, const void *g_src, uint32_t bytes, uint64_t *s_bar) {
    const uint32_t bar = smem_u32(s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s_dst)),
                     "l"(g_src), "r"(bytes), "r"(bar)
                     : "memory");
    }
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar)
        : "memory");
}

// Shared memory: [prog (+aux)] [mbarrier, padded to 16] [dstack] [pstack]
__host__ __device__ inline uint32_t smem_stage_bytes(const ProgView &pv) { return pv.prog_bytes + (pv.stage_aux ? pv.aux_bytes : 0u); }
template <int P>
__host__ __device__ inline uint32_t smem_total_bytes(const ProgView &pv, int threads) {
    return smem_stage_bytes(pv) + 16u + (uint32_t)threads * P * 4u * (pv.dslots + 3u * pv.pslots);
}

template <int P, class Gen>
__global__ void __launch_bounds__(kThreads) k_eval(ProgView pv, Gen gen) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t stage = smem_stage_bytes(pv);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + stage);
    bulk_stage(smem, pv.g_prog, stage, bar);
    const uint4 *prog = reinterpret_cast<const uint4 *>(smem);
    const float4 *aux = pv.stage_aux ? reinterpret_cast<const float4 *>(smem + pv.prog_bytes)
                                     : reinterpret_cast<const float4 *>(reinterpret_cast<const uint8_t *>(pv.g_prog) + pv.prog_bytes);
    float *dstk = reinterpret_cast<float *>(smem + stage + 16u) + threadIdx.x;
    float *pstk = dstk + (size_t)pv.dslots * P * blockDim.x;

    Machine<P> m;
    const uint64_t nwork = gen.work_items();
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwork; w += (uint64_t)gridDim.x * blockDim.x) {
        m.init(dstk, pstk, blockDim.x);
        if (!gen.load(w, m.px, m.py, m.pz)) continue;
        run_program<P>(m, prog, aux);
        gen.store(w, m.top);
    }
}

// ---------------------------------------------------------------------------------------------- generators
// gleval.SDF3.Evaluate on an AoS float3 list (gleval/gleval.go:15-24): 4 points per thread, 3x float4 loads.
struct GenPoints3 {
    const float *pos; float *dist; uint64_t n; int vec;  // vec: both pointers 16-byte aligned
    __device__ uint64_t work_items() const { return (n + 3) / 4; }
    __device__ bool load(uint64_t w, float (&x)[4], float (&y)[4], float (&z)[4]) const {
        const uint64_t i0 = w * 4;
        if (vec && i0 + 4 <= n) {
            const float4 *p4 = reinterpret_cast<const float4 *>(pos) + w * 3;
            const float4 a = __ldg(p4), b = __ldg(p4 + 1), c = __ldg(p4 + 2);
            x[0] = a.x; y[0] = a.y; z[0] = a.z; x[1] = a.w; y[1] = b.x; z[1] = b.y;
            x[2] = b.z; y[2] = b.w; z[2] = c.x; x[3] = c.y; y[3] = c.z; z[3] = c.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint64_t i = i0 + j < n ? i0 + j : n - 1;
                x[j] = __ldg(pos + 3 * i); y[j] = __ldg(pos + 3 * i + 1); z[j] = __ldg(pos + 3 * i + 2);
            }
        }
        return true;
    }
    __device__ void store(uint64_t w, const float (&d)[4]) const {
        const uint64_t i0 = w * 4;
        if (vec && i0 + 4 <= n) {
            reinterpret_cast<float4 *>(dist)[w] = make_float4(d[0], d[1], d[2], d[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (i0 + j < n) dist[i0 + j] = d[j];
        }
    }
};
// gleval.SDF2.Evaluate (gleval/gleval.go:28-37): AoS float2, 2x float4 loads per 4 points.
struct GenPoints2 {
    const float *pos; float *dist; uint64_t n; int vec;
    __device__ uint64_t work_items() const { return (n + 3) / 4; }
    __device__ bool load(uint64_t w, float (&x)[4], float (&y)[4], float (&z)[4]) const {
        const uint64_t i0 = w * 4;
        if (vec && i0 + 4 <= n) {
            const float4 *p4 = reinterpret_cast<const float4 *>(pos) + w * 2;
            const float4 a = __ldg(p4), b = __ldg(p4 + 1);
            x[0] = a.x; y[0] = a.y; x[1] = a.z; y[1] = a.w; x[2] = b.x; y[2] = b.y; x[3] = b.z; y[3] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint64_t i = i0 + j < n ? i0 + j : n - 1;
                x[j] = __ldg(pos + 2 * i); y[j] = __ldg(pos + 2 * i + 1);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) z[j] = 0.f;
        return true;
    }
    __device__ void store(uint64_t w, const float (&d)[4]) const {
        const uint64_t i0 = w * 4;
        if (vec && i0 + 4 <= n) {
            reinterpret_cast<float4 *>(dist)[w] = make_float4(d[0], d[1], d[2], d[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (i0 + j < n) dist[i0 + j] = d[j];
        }
    }
};

// Lattice description shared by the grid / mesher kernels.
struct Lat {
    float ox, oy, oz, res;
    int nx, ny, nz;    // cells
    int k0, nk;        // corner planes [k0, k0+nk) handled
    int nqx;           // quads (4 corners) per row = ceil((nx+1)/4)
    int pitch;         // floats per stored row
    int vec;           // rows 16-byte aligned -> float4 stores
};
// FlatRenderer.evalKRange (glrender/flatrenderer.go:146-182): positions origin + float32(i)*res, x fastest.
// list==nullptr: every quad of the slab; else the compacted quad ids produced by k_compact_quads.
struct GenGrid {
    Lat L; float *dist; const uint32_t *list; const uint32_t *count;
    __device__ uint64_t work_items() const { return list ? (uint64_t)*count : (uint64_t)L.nqx * (L.ny + 1) * L.nk; }
    __device__ void decode(uint64_t w, int &m, int &j, int &k) const {
        uint64_t q = list ? (uint64_t)list[w] : w;
        m = (int)(q % L.nqx); q /= L.nqx;
        j = (int)(q % (L.ny + 1));
        k = (int)(q / (L.ny + 1));
    }
    __device__ bool load(uint64_t w, float (&x)[4], float (&y)[4], float (&z)[4]) const {
        int m, j, k;
        decode(w, m, j, k);
        const float yy = L.oy + (float)j * L.res, zz = L.oz + (float)(L.k0 + k) * L.res;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int i = min(4 * m + t, L.nx);
            x[t] = L.ox + (float)i * L.res; y[t] = yy; z[t] = zz;
        }
        return true;
    }
    __device__ void store(uint64_t w, const float (&d)[4]) const {
        int m, j, k;
        decode(w, m, j, k);
        float *row = dist + ((size_t)k * (L.ny + 1) + j) * L.pitch + 4 * m;
        if (L.vec && 4 * m + 3 < L.pitch) {
            *reinterpret_cast<float4 *>(row) = make_float4(d[0], d[1], d[2], d[3]);
        } else {
#pragma unroll
            for (int t = 0; t < 4; t++) if (4 * m + t <= L.nx) row[t] = d[t];
        }
    }
};

// Octree prune (glrender/octreerenderer.go:180-191, 240-284): evaluate the centre of every level-3 cube (4 cells
// wide) of the slab; keep it iff |d| < size*sqrt3/2. One byte per block, x fastest.
struct GenCenters {
    float ox, oy, oz, res; int nbx, nby, nbz, bz0; float half, maxDist; uint8_t *mask;
    __device__ uint64_t work_items() const { return (uint64_t)((nbx + 3) / 4) * nby * nbz; }
    __device__ void decode(uint64_t w, int &bq, int &by, int &bz) const {
        const int nq = (nbx + 3) / 4;
        bq = (int)(w % nq); w /= nq;
        by = (int)(w % nby); bz = (int)(w / nby);
    }
    __device__ bool load(uint64_t w, float (&x)[4], float (&y)[4], float (&z)[4]) const {
        int bq, by, bz;
        decode(w, bq, by, bz);
        const float yy = (oy + (float)(4 * by) * res) + half, zz = (oz + (float)(4 * (bz0 + bz)) * res) + half;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int bx = min(4 * bq + t, nbx - 1);
            x[t] = (ox + (float)(4 * bx) * res) + half; y[t] = yy; z[t] = zz;
        }
        return true;
    }
    __device__ void store(uint64_t w, const float (&d)[4]) const {
        int bq, by, bz;
        decode(w, bq, by, bz);
        uint8_t *row = mask + ((size_t)bz * nby + by) * nbx;
#pragma unroll
        for (int t = 0; t < 4; t++)
            if (4 * bq + t < nbx) row[4 * bq + t] = fabsf(d[t]) >= maxDist ? 0 : 1;
    }
};

// ImageRendererSDF2.Render positions (glrender/image.go:85-105).
struct GenImage {
    float xmin, ymax, dx, dy; int w, h; float *dist;
    __device__ uint64_t work_items() const { return (uint64_t)((w + 3) / 4) * h; }
    __device__ bool load(uint64_t wi, float (&x)[4], float (&y)[4], float (&z)[4]) const {
        const int nq = (w + 3) / 4;
        const int q = (int)(wi % nq), j = (int)(wi / nq);
        const float yy = ymax - (float)j * dy;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int i = min(4 * q + t, w - 1);
            x[t] = (float)i * dx + xmin; y[t] = yy; z[t] = 0.f;
        }
        return true;
    }
    __device__ void store(uint64_t wi, const float (&d)[4]) const {
        const int nq = (w + 3) / 4;
        const int q = (int)(wi % nq), j = (int)(wi / nq);
        float *row = dist + (size_t)j * w;
        if ((w & 3) == 0) {
            *reinterpret_cast<float4 *>(row + 4 * q) = make_float4(d[0], d[1], d[2], d[3]);
        } else {
#pragma unroll
            for (int t = 0; t < 4; t++) if (4 * q + t < w) row[4 * q + t] = d[t];
        }
    }
};

// ---------------------------------------------------------------------------------------------- prune -> quad list
struct MeshDims {
    int nx, ny, nz;          // cells of the whole lattice
    int cz0, cz1;            // slab of cells
    int nbx, nby, nbz, bz0;  // 4-cell blocks covering the slab
    int nqx;                 // quads per corner row
    int pitch;               // grid row pitch (floats)
    int nsx;                 // 32-cell segments per cell row
};

__device__ __forceinline__ bool block_on(const uint8_t *mask, const MeshDims &D, int bx, int by, int bz) {
    if (bx < 0 || by < 0 || bx >= D.nbx || by >= D.nby) return false;
    const int lz = bz - D.bz0;
    if (lz < 0 || lz >= D.nbz) return false;
    return mask[((size_t)lz * D.nby + by) * D.nbx + bx] != 0;
}

// One thread per lattice quad (m,j,k) of the slab's corner planes: it must be evaluated iff some kept block owns a
// cell that touches one of its 4 corners. Survivors are appended warp-aggregated (ballot + one atomicAdd per warp).
__global__ void __launch_bounds__(kThreads) k_compact_quads(MeshDims D, const uint8_t *__restrict__ mask, uint32_t *__restrict__ list,
                                                           uint32_t *__restrict__ count) {
    const uint64_t nq = (uint64_t)D.nqx * (D.ny + 1) * (D.cz1 - D.cz0 + 1);
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < nq; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t q = base + threadIdx.x;
        bool need = false;
        if (q < nq) {
            uint64_t t = q;
            const int m = (int)(t % D.nqx); t /= D.nqx;
            const int j = (int)(t % (D.ny + 1));
            const int k = D.cz0 + (int)(t / (D.ny + 1));
            // cells touching corner plane k inside the slab: cz = k-1 (if >= cz0) and cz = k (if < cz1)
#pragma unroll
            for (int dz = -1; dz <= 0; dz++) {
                const int cz = k + dz;
                if (cz < D.cz0 || cz >= D.cz1) continue;
#pragma unroll
                for (int dy = -1; dy <= 0; dy++) {
                    const int cy = j + dy;
                    if (cy < 0 || cy >= D.ny) continue;
                    // cells cx in [4m-1, 4m+3] -> blocks m-1 (via cx=4m-1) and m
                    if (4 * m - 1 >= 0 && block_on(mask, D, m - 1, cy >> 2, cz >> 2)) need = true;
                    if (4 * m < D.nx && block_on(mask, D, m, cy >> 2, cz >> 2)) need = true;
                }
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, need);
        if (bal) {
            const int lane = threadIdx.x & 31;
            uint32_t wbase = 0;
            if (lane == 0) wbase = atomicAdd(count, (uint32_t)__popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (need) list[wbase + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)q;
        }
    }
}

// ---------------------------------------------------------------------------------------------- marching cubes
struct MCArgs {
    MeshDims D;
    float ox, oy, oz, res, cubeDiag;
    const float *grid;      // slab corner planes, plane 0 = corner plane cz0
    const uint8_t *mask;    // nullptr = FlatRenderer semantics (no prune)
    uint32_t *segcount;     // per segment triangle count (count pass) / exclusive offsets (emit pass)
    float *tris;            // 9 floats per triangle
    uint64_t tri_capacity;
    uint8_t *cases;         // optional nx*ny*(cz1-cz0) bytes
    uint32_t *overflow;     // set to 1 if a triangle did not fit
};

// marchcubes.go:39-44 + flatrenderer.go:215-233: corner order (0,0,0)(1,0,0)(1,1,0)(0,1,0)(0,0,1)(1,0,1)(1,1,1)(0,1,1)
__device__ __forceinline__ int mc_classify(const MCArgs &A, int cx, int cy, int cz, float (&v)[8]) {
    const MeshDims &D = A.D;
    if (A.mask && !block_on(A.mask, D, cx >> 2, cy >> 2, cz >> 2)) return 0;
    const size_t sy = (size_t)D.pitch, sz = sy * (D.ny + 1);
    const float *g = A.grid + (size_t)(cz - D.cz0) * sz + (size_t)cy * sy + cx;
    v[0] = __ldg(g);
    if (fabsf(v[0]) > A.cubeDiag) return 0;  // flatrenderer.go:218-220
    v[1] = __ldg(g + 1); v[2] = __ldg(g + 1 + sy); v[3] = __ldg(g + sy);
    v[4] = __ldg(g + sz); v[5] = __ldg(g + 1 + sz); v[6] = __ldg(g + 1 + sy + sz); v[7] = __ldg(g + sy + sz);
    int index = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) if (v[i] < 0.f) index |= 1 << i;
    return index;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Pass 1: triangles per 32-cell row segment. One warp per segment.
__global__ void __launch_bounds__(kThreads) k_mc_count(MCArgs A) {
    __shared__ uint8_t s_ntri[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = c_mc_ntri[i];
    __syncthreads();
    const MeshDims &D = A.D;
    const uint64_t nseg = (uint64_t)D.nsx * D.ny * (D.cz1 - D.cz0);
    const int lane = threadIdx.x & 31;
    const uint64_t wpg = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t s = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < nseg; s += wpg) {
        uint64_t t = s;
        const int sx = (int)(t % D.nsx); t /= D.nsx;
        const int cy = (int)(t % D.ny);
        const int cz = D.cz0 + (int)(t / D.ny);
        const int cx = sx * 32 + lane;
        int index = 0;
        float v[8];
        if (cx < D.nx) index = mc_classify(A, cx, cy, cz, v);
        if (A.cases && cx < D.nx) A.cases[((size_t)(cz - D.cz0) * D.ny + cy) * D.nx + cx] = (uint8_t)index;
        uint32_t n = s_ntri[index];
        const uint32_t incl = warp_incl_scan(n);
        if (lane == 31) A.segcount[s] = incl;
    }
}

// marchcubes.go:76-98
__device__ __forceinline__ float3 mc_interp(float3 p1, float3 p2, float v1, float v2) {
    const float eps = 1e-12f;
    const bool c1 = fabsf(0.f - v1) < eps, c2 = fabsf(0.f - v2) < eps;
    if (c1 && !c2) return p1;
    if (c2 && !c1) return p2;
    float t = 0.5f;
    if (!c1 || !c2) t = (0.f - v1) / (v2 - v1);
    return make_float3(p1.x + t * (p2.x - p1.x), p1.y + t * (p2.y - p1.y), p1.z + t * (p2.z - p1.z));
}

// Pass 2: emit. segcount now holds exclusive offsets; the warp inclusive scan places each cell's triangles.
__global__ void __launch_bounds__(kThreads) k_mc_emit(MCArgs A) {
    __shared__ uint8_t s_ntri[256];
    __shared__ int8_t s_tris[256 * 16];
    __shared__ uint16_t s_edges[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_ntri[i] = c_mc_ntri[i]; s_edges[i] = c_mc_edges[i]; }
    for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) s_tris[i] = c_mc_tris[i];
    __syncthreads();
    const MeshDims &D = A.D;
    const uint64_t nseg = (uint64_t)D.nsx * D.ny * (D.cz1 - D.cz0);
    const int lane = threadIdx.x & 31;
    const uint64_t wpg = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t s = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < nseg; s += wpg) {
        uint64_t t = s;
        const int sx = (int)(t % D.nsx); t /= D.nsx;
        const int cy = (int)(t % D.ny);
        const int cz = D.cz0 + (int)(t / D.ny);
        const int cx = sx * 32 + lane;
        int index = 0;
        float v[8];
        if (cx < D.nx) index = mc_classify(A, cx, cy, cz, v);
        const uint32_t n = s_ntri[index];
        const uint32_t incl = warp_incl_scan(n);
        if (n == 0) continue;
        uint64_t o = (uint64_t)A.segcount[s] + (incl - n);
        // corner positions, flatrenderer.go:235-247
        const float r = A.res;
        const float x0 = A.ox + (float)cx * r, y0 = A.oy + (float)cy * r, z0 = A.oz + (float)cz * r;
        const float x1 = x0 + r, y1 = y0 + r, z1 = z0 + r;
        const float3 p[8] = {{x0, y0, z0}, {x1, y0, z0}, {x1, y1, z0}, {x0, y1, z0}, {x0, y0, z1}, {x1, y0, z1}, {x1, y1, z1}, {x0, y1, z1}};
        const uint32_t edges = s_edges[index];
        float3 pts[12];
        // edge -> corner pairs (marchcubes.go:101-114), compile-time so p[]/v[] stay in registers
        constexpr int PA[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
        constexpr int PB[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
#pragma unroll
        for (int e = 0; e < 12; e++) {
            if (edges & (1u << e)) pts[e] = mc_interp(p[PA[e]], p[PB[e]], v[PA[e]], v[PB[e]]);
        }
        const int8_t *tb = s_tris + 16 * index;
        for (uint32_t k = 0; k < n; k++, o++) {
            if (o >= A.tri_capacity) { *A.overflow = 1u; break; }
            // marchcubes.go:64-68: (points[t+2], points[t+1], points[t])
            const float3 a = pts[tb[3 * k + 2]], b = pts[tb[3 * k + 1]], c = pts[tb[3 * k]];
            float *dst = A.tris + 9 * o;
            dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = b.x; dst[4] = b.y; dst[5] = b.z; dst[6] = c.x; dst[7] = c.y; dst[8] = c.z;
        }
    }
}

// ---------------------------------------------------------------------------------------------- exclusive scan
constexpr int kScanItems = 4;  // per thread; 1024 per block
__global__ void __launch_bounds__(kThreads) k_scan_reduce(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ blocksum) {
    __shared__ uint32_t s_w[kThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * (kThreads * kScanItems);
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        const uint64_t idx = base + (uint64_t)threadIdx.x * kScanItems + i;
        if (idx < n) v += in[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < kThreads / 32; i++) t += s_w[i];
        blocksum[blockIdx.x] = t;
    }
}
// single CTA: exclusive scan of block sums in place; total (64-bit) to *total
__global__ void __launch_bounds__(1024) k_scan_blocksums(uint32_t *__restrict__ blocksum, uint32_t nblocks, unsigned long long *__restrict__ total) {
    __shared__ uint32_t s_w[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0ull;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += 1024) {
        const uint32_t idx = base + threadIdx.x;
        const uint32_t v = idx < nblocks ? blocksum[idx] : 0u;
        uint32_t incl = warp_incl_scan(v);
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = s_w[threadIdx.x];
            uint32_t wi = warp_incl_scan(w);
            s_w[threadIdx.x] = wi - w;
        }
        __syncthreads();
        const unsigned long long carry = s_carry;
        const uint32_t excl = incl - v + s_w[threadIdx.x >> 5];
        if (idx < nblocks) blocksum[idx] = (uint32_t)(carry + excl);
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}
__global__ void __launch_bounds__(kThreads) k_scan_apply(uint32_t *__restrict__ data, uint64_t n, const uint32_t *__restrict__ blocksum) {
    __shared__ uint32_t s_w[kThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * (kThreads * kScanItems) + (uint64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) { v[i] = base + i < n ? data[base + i] : 0u; sum += v[i]; }
    const uint32_t incl = warp_incl_scan(sum);
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = threadIdx.x < kThreads / 32 ? s_w[threadIdx.x] : 0u;
        uint32_t wi = warp_incl_scan(w);
        if (threadIdx.x < kThreads / 32) s_w[threadIdx.x] = wi - w;
    }
    __syncthreads();
    uint32_t run = blocksum[blockIdx.x] + s_w[threadIdx.x >> 5] + (incl - sum);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
}

// ---------------------------------------------------------------------------------------------- STL
// glrender/stl.go:33-61: 50-byte records (unit normal, 3 vertices, u16 0). Records are built in shared memory and
// written out as aligned 32-bit words; `out` points at the first record and must be 4-byte aligned.
__global__ void __launch_bounds__(kThreads) k_stl_pack(const float *__restrict__ tri9, uint64_t ntri, uint8_t *__restrict__ out) {
    __shared__ __align__(16) uint8_t s_rec[kThreads * 50];
    for (uint64_t base = (uint64_t)blockIdx.x * kThreads; base < ntri; base += (uint64_t)gridDim.x * kThreads) {
        const uint64_t i = base + threadIdx.x;
        if (i < ntri) {
            const float *t = tri9 + 9 * i;
            float v[9];
#pragma unroll
            for (int k = 0; k < 9; k++) v[k] = __ldg(t + k);
            // ms3.Triangle.Normal = Cross(t1-t0, t2-t1); ms3.Unit = Scale(1/Norm(n), n)
            const float s1x = v[3] - v[0], s1y = v[4] - v[1], s1z = v[5] - v[2];
            const float s2x = v[6] - v[3], s2y = v[7] - v[4], s2z = v[8] - v[5];
            float nx = s1y * s2z - s1z * s2y, ny = s1z * s2x - s1x * s2z, nz = s1x * s2y - s1y * s2x;
            if (nx == 0.f && ny == 0.f && nz == 0.f) {
                nx = ny = nz = __int_as_float(0x7fc00000);
            } else {
                const float inv = 1.f / m32::norm3(nx, ny, nz);
                nx *= inv; ny *= inv; nz *= inv;
            }
            uint16_t *r = reinterpret_cast<uint16_t *>(s_rec + threadIdx.x * 50);
            const float f[12] = {nx, ny, nz, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]};
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const uint32_t u = __float_as_uint(f[k]);
                r[2 * k] = (uint16_t)(u & 0xffffu);
                r[2 * k + 1] = (uint16_t)(u >> 16);
            }
            r[24] = 0;
        }
        __syncthreads();
        const uint64_t nvalid = ntri - base < (uint64_t)kThreads ? ntri - base : (uint64_t)kThreads;
        const uint32_t nbytes = (uint32_t)nvalid * 50u;
        uint8_t *dst = out + base * 50;
        // base*50 is a multiple of 4 because base is a multiple of 256
        const uint32_t nwords = nbytes / 4u;
        for (uint32_t w = threadIdx.x; w < nwords; w += blockDim.x)
            reinterpret_cast<uint32_t *>(dst)[w] = reinterpret_cast<const uint32_t *>(s_rec)[w];
        for (uint32_t b = nwords * 4u + threadIdx.x; b < nbytes; b += blockDim.x) dst[b] = s_rec[b];
        __syncthreads();
    }
}

}  // namespace gsdfk
