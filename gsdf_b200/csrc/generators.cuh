// generators.cuh -- what every translation unit of the device library shares: the kernel-side view of a node program,
// programmatic-dependent-launch helpers, the bulk-async program staging, the lattice description and the generator
// functors that feed the interpreter kernel (eval_kernels.cuh) with positions and take its distances.
#pragma once
#ifdef __CUDACC_RTC__  // run-time compilation of a specialised kernel (jit.cu): no system headers, no image generator
#include "rtc_types.cuh"
#else
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "../../include/gsdf_program.h"
#ifndef __CUDACC_RTC__
#include "colormap.cuh"
#endif

namespace gsdfk {

#ifndef GSDF_THREADS
#define GSDF_THREADS 256
#endif
constexpr int kThreads = GSDF_THREADS;   // CTA size of the MC / scan / STL kernels and default of k_eval
// Measured on B200 (scripts/ab_eval.py): CTAs whose warps are kept on the same opcode body by a barrier per instruction
// (GSDF_LOCKSTEP) cut instruction-fetch stalls: -13 % (flange) / -14 % (knurled) evaluate time versus free-running
// 256-thread CTAs. CTA size (scripts/gpu_r2_ab_cta.sh, graph replays, flange@400 / bolt@400 / knurled@500 evaluate time):
// 512 threads x 2 CTAs per SM (64 registers) 77 / 77 / 449 us; 384 x 3 (56 registers, no spill: 36 warps per SM and three
// independent lockstep convoys) 70 / 72 / 442 us; 256 x 4 69 / 66 / 470; 128 x 8 67 / 63 / 508 (more convoys, but the
// many-opcode knurled tree thrashes the instruction cache); register caps that spill (48, 40) lose everywhere.
#ifndef GSDF_EVAL_THREADS
#define GSDF_EVAL_THREADS 384
#endif
constexpr int kEvalThreads = GSDF_EVAL_THREADS;  // CTA size of the interpreter kernel
#ifndef GSDF_EVAL_MINB
#define GSDF_EVAL_MINB 3  // resident CTAs per SM the interpreter kernel is compiled for (register cap 65536 / threads / MINB)
#endif
constexpr int kImgTileRows = kEvalThreads / 32;  // image work items: tiles of 32 quads x kImgTileRows rows = one CTA tile

struct ProgView {
    const uint4 *g_prog;     // device: program chunks followed by aux (16-byte aligned)
    uint32_t prog_bytes;     // bytes of chunks
    uint32_t aux_bytes;      // bytes of aux that follow the chunks
    uint32_t stage_aux;      // 1: aux is staged to smem with the program; 0: read from global
    uint32_t dslots, pslots; // stack slots
    uint32_t *sched;         // [0] next tile, [1] finished CTAs (self-resetting work counter)
    unsigned long long *stamp;  // optional: receives %globaltimer when the kernel's first CTA starts its work (stage timing inside graphs)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Programmatic dependent launch (PTX griddepcontrol). The kernels of one render form a chain in which each consumes
// what its predecessor wrote. Every kernel of the chain (a) lets its successor's CTAs become resident as soon as all of
// its own CTAs have started (pdl_trigger) and (b) does whatever does not depend on the predecessor -- staging the node
// program, loading tables, initialising mbarriers -- before pdl_wait(), which returns once the predecessor grid has
// completed and its writes are visible. Because every kernel waits before it touches chain data, completion is
// transitive along the chain. Launched without the programmatic-serialization attribute both are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Stage stamps: every kernel of a render stores %globaltimer (ns) when its first CTA passes pdl_wait(), i.e. when its
// predecessor has completed; differences of consecutive stamps are the stage times INSIDE a CUDA-graph replay, where
// events cannot be recorded between the kernels without breaking the programmatic edges.
__device__ __forceinline__ void stage_stamp(unsigned long long *slot) {
    if (slot && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        *slot = t;
    }
}

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

// Shared memory: [prog (+aux)] [mbarrier + tile slot, 16 bytes] [dstack] [pstack]
__host__ __device__ inline uint32_t smem_stage_bytes(const ProgView &pv) { return pv.prog_bytes + (pv.stage_aux ? pv.aux_bytes : 0u); }
// Radius cache (gsdf_program.h, "Radius reuse"; absent from -DGSDF_NO_RXY builds): one float per point behind the stacks.
#ifdef GSDF_RXY
constexpr uint32_t kRxySlots = 1u;
#else
constexpr uint32_t kRxySlots = 0u;
#endif
template <int P>
__host__ __device__ inline uint32_t smem_total_bytes(const ProgView &pv, int threads) {
    return smem_stage_bytes(pv) + 16u + (uint32_t)threads * P * 4u * (pv.dslots + 3u * pv.pslots + kRxySlots);
}


// ---------------------------------------------------------------------------------------------- generators
// gleval.SDF3.Evaluate on an AoS float3 list (gleval/gleval.go:15-24): 4 points per thread, 3x float4 loads.
struct GenPoints3 {
    static constexpr bool kTileSkip = false;
    const float *pos; float *dist; uint64_t n; int vec;  // vec: both pointers 16-byte aligned
    __device__ uint64_t work_items() const { return (n + 3) / 4; }
    __device__ void load(uint64_t w, float (&x)[4], float (&y)[4], float (&z)[4]) const {
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
    static constexpr bool kTileSkip = false;
    const float *pos; float *dist; uint64_t n; int vec;
    __device__ uint64_t work_items() const { return (n + 3) / 4; }
    __device__ void load(uint64_t w, float (&x)[4], float (&y)[4], float (&z)[4]) const {
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

// n / d for n < 2^31 and a run-time divisor as __umulhi(n, mul) >> shr (mul == 0: divisor 1); a 32-bit hardware-less division
// by a run-time divisor is ~18-35 instructions.
__host__ __device__ inline void fastdiv_init(uint32_t d, uint32_t &mul, uint32_t &shr) {
    if (d <= 1u) { mul = 0u; shr = 0u; return; }
    uint32_t lg = 0;
    while ((1ull << lg) < d) lg++;          // ceil(log2 d)
    const uint32_t p = 31u + lg;
    mul = (uint32_t)(((1ull << p) + d - 1u) / d);
    shr = p - 32u;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, uint32_t mul, uint32_t shr) { return mul ? __umulhi(n, mul) >> shr : n; }


// Lattice description shared by the grid / mesher kernels.
struct Lat {
    float ox, oy, oz, res;
    int nx, ny, nz;    // cells
    int k0, nk;        // corner planes [k0, k0+nk) handled
    int nqx;           // quads (4 corners) per row = ceil((nx+1)/4)
    int pitch;         // floats per stored row
    int vec;           // rows 16-byte aligned -> float4 stores
    // quad id -> (m, j, k) by multiply-shift (fastdiv) when the slab has fewer than 2^31 quads (fdiv != 0): decode runs twice per
    // work item (load, store), two divisions each -- 10 % of the instructions of a one-corner-per-thread evaluation of the flange
    uint32_t fdiv, nqx_mul, nqx_shr, nyp_mul, nyp_shr;
    // hq != 0 (two corners per thread only): the work list holds HALF-quads, 2 * quad + half. A quad whose own block is not
    // kept is listed only because its first corner is the last corner column of the kept block on its left: its second half is
    // never read by anybody and is not listed (5.5 % of the flange's lattice evaluations, 9 % of the knurled cylinder's)
    int hq;
};
// FlatRenderer.evalKRange (glrender/flatrenderer.go:146-182): positions origin + float32(i)*res, x fastest.
// Work unit = a quad of 4 consecutive corners of one lattice row, split over 4/P threads.
// list==nullptr: every quad of the slab; else the compacted quad ids produced by k_compact_quads.
template <int P>
struct GenGrid {
    static constexpr bool kTileSkip = false;
    static_assert(P == 1 || P == 2 || P == 4, "P must divide 4");
    Lat L; float *dist; const uint32_t *list; const uint32_t *count;
    __device__ uint64_t work_items() const {
        if (P == 2 && list && L.hq) return (uint64_t)*count;
        return (list ? (uint64_t)*count : (uint64_t)L.nqx * (L.ny + 1) * L.nk) * (4 / P);
    }
    __device__ void decode(uint64_t w, int &i0, int &j, int &k) const {
        uint32_t sub = (uint32_t)(w % (4 / P));
        uint32_t q;
        if (P == 2 && list && L.hq) { const uint32_t h = list[w]; q = h >> 1; sub = h & 1u; }
        else q = list ? list[w / (4 / P)] : (uint32_t)(w / (4 / P));
        uint32_t m;
        if (L.fdiv) {
            const uint32_t r = fastdiv(q, L.nqx_mul, L.nqx_shr);
            m = q - r * (uint32_t)L.nqx;
            k = (int)fastdiv(r, L.nyp_mul, L.nyp_shr);
            j = (int)(r - (uint32_t)k * (uint32_t)(L.ny + 1));
        } else {
            m = q % (uint32_t)L.nqx; q /= (uint32_t)L.nqx;
            j = (int)(q % (uint32_t)(L.ny + 1));
            k = (int)(q / (uint32_t)(L.ny + 1));
        }
        i0 = (int)(4 * m + P * sub);
    }
    __device__ void load(uint64_t w, float (&x)[P], float (&y)[P], float (&z)[P]) const {
        int i0, j, k;
        decode(w, i0, j, k);
        const float yy = L.oy + (float)j * L.res, zz = L.oz + (float)(L.k0 + k) * L.res;
#pragma unroll
        for (int t = 0; t < P; t++) {
            const int i = min(i0 + t, L.nx);
            x[t] = L.ox + (float)i * L.res; y[t] = yy; z[t] = zz;
        }
    }
    __device__ void store(uint64_t w, const float (&d)[P]) const {
        int i0, j, k;
        decode(w, i0, j, k);
        float *row = dist + ((size_t)k * (L.ny + 1) + j) * L.pitch + i0;
        if (L.vec) {  // pitch is a multiple of 4 and covers every quad: no bounds check needed
            if (P == 4) *reinterpret_cast<float4 *>(row) = make_float4(d[0], d[P > 1 ? 1 : 0], d[P > 2 ? 2 : 0], d[P > 3 ? 3 : 0]);
            else if (P == 2) *reinterpret_cast<float2 *>(row) = make_float2(d[0], d[P > 1 ? 1 : 0]);
            else row[0] = d[0];
        } else {
#pragma unroll
            for (int t = 0; t < P; t++) if (i0 + t <= L.nx) row[t] = d[t];
        }
    }
};

// Octree prune (glrender/octreerenderer.go:180-191, 240-284), one level of the coarse-to-fine plan: evaluate the centre of
// every level-L cube (w = 2^(L-1) cells wide; level 3 = 4 cells) of the slab whose parent cube survived the previous,
// coarser level, and keep it iff |d| < margin * size*sqrt3/2 (margin 1 = the reference's literal rule). Cubes of every
// level are aligned to the lattice origin like the reference's octree cubes (ms3.Octree.CubeOrigin), so a cube's
// verdict depends on its own centre and its ancestors' only -- a Z-slab computes the same bits as the whole lattice.
// One cube per thread: the pass is small and latency bound, so it wants threads, not per-thread ILP. Work items are cube
// rows padded to whole warps (32*nwx per row), so a warp's 32 verdicts are exactly one word of the level's bit rows: the
// sink writes the word with one ballot. Tiles whose cubes all have pruned parents are skipped CTA-uniformly.
struct PruneLevel {
    int w;                    // cube width in cells
    int ncx, ncy, ncz, cz0;   // cubes covering the slab; cz0 = first cube layer (global cube coordinates)
    int nwx;                  // 32-bit words per cube row of the bit mask
    float half, maxDist;      // size/2 and margin * size * sqrt3/2
    uint32_t *bits;           // [ncz][ncy][nwx]
    // work item -> (cx, cy, cz) by multiply-shift when the level has fewer than 2^31 padded cubes (fdiv != 0): GenCenters decodes
    // three times per cube (dead, load, store), and w / rowlen on a 64-bit work item is the most expensive division there is
    uint32_t fdiv, row_mul, row_shr, ncy_mul, ncy_shr;
};
struct GenCenters {
    static constexpr bool kTileSkip = true;
    float ox, oy, oz, res;
    PruneLevel L;             // the level being evaluated
    PruneLevel P;             // its parent level (P.bits == nullptr: L is the top level, every cube is live)
    int shift;                // log2(P.w / L.w)
    uint32_t *kept;           // += surviving cubes of this level (TotalPruned bookkeeping; may be nullptr)
    uint32_t *evals;          // += centres actually evaluated
    __device__ uint64_t work_items() const { return (uint64_t)L.nwx * 32u * L.ncy * L.ncz; }
    __device__ void decode(uint64_t w, int &cx, int &cy, int &cz) const {
        const uint32_t rowlen = (uint32_t)L.nwx * 32u;
        if (L.fdiv) {
            const uint32_t w32 = (uint32_t)w;
            const uint32_t row = fastdiv(w32, L.row_mul, L.row_shr);
            cx = (int)(w32 - row * rowlen);
            cz = (int)fastdiv(row, L.ncy_mul, L.ncy_shr);
            cy = (int)(row - (uint32_t)cz * (uint32_t)L.ncy);
            return;
        }
        const uint32_t row = (uint32_t)(w / rowlen);
        cx = (int)(w - (uint64_t)row * rowlen);
        cy = (int)(row % (uint32_t)L.ncy);
        cz = (int)(row / (uint32_t)L.ncy);
    }
    __device__ bool live(int cx, int cy, int cz) const {
        if (cx >= L.ncx) return false;
        if (!P.bits) return true;
        const int px = cx >> shift, py = cy >> shift, pz = ((L.cz0 + cz) >> shift) - P.cz0;
        return (P.bits[((size_t)pz * P.ncy + py) * P.nwx + (px >> 5)] >> (px & 31)) & 1u;
    }
    __device__ bool dead(uint64_t w) const {
        int cx, cy, cz;
        decode(w, cx, cy, cz);
        return !live(cx, cy, cz);
    }
    __device__ void store_dead(uint64_t w) const {
        if ((threadIdx.x & 31) == 0) L.bits[w >> 5] = 0u;
    }
    __device__ void load(uint64_t w, float (&x)[1], float (&y)[1], float (&z)[1]) const {
        int cx, cy, cz;
        decode(w, cx, cy, cz);
        cx = min(cx, L.ncx - 1);  // padding lanes repeat the row's last cube
        x[0] = (ox + (float)(L.w * cx) * res) + L.half;
        y[0] = (oy + (float)(L.w * cy) * res) + L.half;
        z[0] = (oz + (float)(L.w * (L.cz0 + cz)) * res) + L.half;
    }
    __device__ void store(uint64_t w, const float (&d)[1]) const {
        int cx, cy, cz;
        decode(w, cx, cy, cz);
        const bool alive = live(cx, cy, cz);
        const bool keep = alive && !(fabsf(d[0]) >= L.maxDist);
        const uint32_t word = __ballot_sync(__activemask(), keep);  // work items are multiples of 32: whole warps arrive here
        const uint32_t nlive = (uint32_t)__popc(__ballot_sync(__activemask(), alive));
        if ((threadIdx.x & 31) == 0) {
            L.bits[w >> 5] = word;  // word (w>>5) = row * nwx + cx/32
            if (word && kept) atomicAdd(kept, (uint32_t)__popc(word));
            if (nlive) atomicAdd(evals, nlive);
        }
    }
};

// The last two levels of a plan that ends with level 2 (2-cell cubes), in ONE launch (k_prune_fine, eval_kernels.cuh): a CTA
// evaluates the centres of a tile of level-3 cubes, compacts the survivors in shared memory and evaluates their eight
// children in the same tile loop -- no launch between the two levels, and a child is only ever evaluated by the CTA that
// kept its parent. A level-3 cube none of whose children survives is dropped as well. Outputs: the level-3 bit rows
// (what the marching-cubes stage and its work lists read), one child-mask byte per level-3 cube (bit dx + 2 dy + 4 dz;
// the marching-cubes kernels mask their cells with it: corners of dropped children are never evaluated) and the level-2
// rows the quad list is built from: per level-2 row (2 cz + dz, 2 cy + dy) and per 32 parents two words, E = children with
// dx = 0 and O = children with dx = 1, so that "quad m is touched" is E | O | (O << 1 | carry) without bit interleaving.
struct PruneFine {
    float ox, oy, oz, res;
    PruneLevel L;             // level 3 on this slab (bits: output)
    PruneLevel P;             // its parent level (P.bits == nullptr: none)
    int shift;                // log2(P.w / L.w)
    int nx, ny, nz;           // cells of the whole lattice: a 2-cell cube exists iff its first cell does
    float half2, maxDist2;    // level 2: size/2 and margin * size * sqrt3/2
    uint32_t *bits2;          // [2 L.ncz][2 L.ncy][2 L.nwx]: words (E, O) per 32 parents
    uint8_t *childmask;       // [L.ncz][L.ncy][L.ncx]
    uint32_t *kept;           // += level-3 cubes that keep at least one child
    uint32_t *kept2;          // += surviving level-2 cubes
    uint32_t *evals;          // += centres evaluated (both levels)
};

#ifndef __CUDACC_RTC__
// ImageRendererSDF2.Render positions (glrender/image.go:85-105). rgba != nullptr: the colour conversion is applied in
// the sink and four RGBA8 pixels leave as one 16-byte store (image.go:112-116 fused); else the distances are stored.
struct GenImage {
    static constexpr bool kTileSkip = false;
    // Work items are grouped into 2-D tiles of 32 quads x kImgTileRows rows (128 x 12 pixels = one 384-thread CTA tile), so that
    // a tile is spatially compact and the CTA-uniform guards (gsdf_program.h) fire; a warp still covers 512 contiguous
    // bytes of one image row.
    float xmin, ymax, dx, dy; int w, h; float *dist; uint32_t *rgba; ColorConv cc;
    __host__ __device__ static uint64_t items_for(int w, int h) {
        return (uint64_t)((((uint32_t)(w + 3) / 4) + 31u) / 32u) * (((uint32_t)h + kImgTileRows - 1u) / kImgTileRows) * (32u * kImgTileRows);
    }
    __device__ uint32_t tiles_x() const { return ((uint32_t)(w + 3) / 4 + 31u) / 32u; }
    __device__ uint64_t work_items() const { return items_for(w, h); }
    __device__ void decode(uint64_t wi, int &q, int &j) const {
        const uint32_t tile = (uint32_t)(wi / (32u * kImgTileRows)), t = (uint32_t)(wi - (uint64_t)tile * (32u * kImgTileRows));
        const uint32_t tx = tile % tiles_x(), ty = tile / tiles_x();
        q = (int)(tx * 32u + (t & 31u));
        j = (int)(ty * kImgTileRows + (t >> 5));
    }
    __device__ void load(uint64_t wi, float (&x)[4], float (&y)[4], float (&z)[4]) const {
        int q, j;
        decode(wi, q, j);
        const float yy = ymax - (float)min(j, h - 1) * dy;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int i = min(4 * q + t, w - 1);
            x[t] = (float)i * dx + xmin; y[t] = yy; z[t] = 0.f;
        }
    }
    __device__ void store(uint64_t wi, const float (&d)[4]) const {
        int q, j;
        decode(wi, q, j);
        if (j >= h || 4 * q >= w) return;
        if (rgba) {
            uint32_t c[4];
#pragma unroll
            for (int t = 0; t < 4; t++) c[t] = color_of(cc, d[t]);
            uint32_t *row = rgba + (size_t)j * w;
            if ((w & 3) == 0) {
                *reinterpret_cast<uint4 *>(row + 4 * q) = make_uint4(c[0], c[1], c[2], c[3]);
            } else {
#pragma unroll
                for (int t = 0; t < 4; t++) if (4 * q + t < w) row[4 * q + t] = c[t];
            }
            return;
        }
        float *row = dist + (size_t)j * w;
        if ((w & 3) == 0) {
            *reinterpret_cast<float4 *>(row + 4 * q) = make_float4(d[0], d[1], d[2], d[3]);
        } else {
#pragma unroll
            for (int t = 0; t < 4; t++) if (4 * q + t < w) row[4 * q + t] = d[t];
        }
    }
};

#endif  // __CUDACC_RTC__

// ---------------------------------------------------------------------------------------------- prune -> quad list
struct MeshDims {
    int nx, ny, nz;          // cells of the whole lattice
    int cz0, cz1;            // slab of cells
    int nbx, nby, nbz, bz0;  // 4-cell blocks covering the slab
    int nqx;                 // quads per corner row
    int pitch;               // grid row pitch (floats)
    int nsx;                 // 32-cell segments per cell row
    int nwx;                 // 32-bit words per block row of the bit mask = ceil(nbx/32)
    // n / nbx and n / nby for n < 2^31 as __umulhi(n, mul) >> shr (mul == 0: divisor 1): the block kernels decode two block
    // ids per block, and a 32-bit division by a run-time divisor is ~35 instructions
    uint32_t nbx_mul, nbx_shr, nby_mul, nby_shr;
};
}  // namespace gsdfk
