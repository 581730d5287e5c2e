// mc_kernels.cuh -- sm_100a kernels of the mesh stage (compiled in mesher.cu).
//
//   k_compact_quads    octree level-3 prune mask -> compacted list of 4-corner lattice quads that still need evaluating
//   k_mc_count/emit    marching cubes. Pass 1 writes a triangle count per 32-cell row segment and a compact list of
//                      non-empty segments; an exclusive scan turns the counts into offsets; pass 2 (warp per listed
//                      segment) places each cell's triangles. Output order is the reference FlatRenderer's (cell index
//                      x fastest, flatrenderer.go:208-212) and fully deterministic.
//   k_scan_*           exclusive scans over segment counts (decoupled look-back, and the three-kernel A/B form).
//   k_stl_pack         glrender/stl.go:33-61 record packing.
#pragma once
#include "generators.cuh"
#include "math32.cuh"
#include "mc_tables.cuh"

namespace gsdfk {

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__device__ __forceinline__ uint32_t bit_at(const uint32_t *row, int b) { return b < 0 ? 0u : (row[b >> 5] >> (b & 31)) & 1u; }

// Quad list. A quad (4 consecutive corners of corner row (j,k) starting at 4m) must be evaluated iff some kept block owns a
// cell touching one of its corners. Such cells have cy in {j-1,j}, cz in {k-1,k} (inside the slab) and cx in [4m-1, 4m+3],
// i.e. blocks m-1 and m of up to four block rows. Work item = one 32-quad WORD of one corner row (lane per item, so a
// 71-quad row costs 3 lanes, not a warp): need-word = word | word<<1 | carry of the previous word over the touching block
// rows; the survivors of a warp's 32 words are appended with one atomicAdd per warp and expanded cooperatively (lane q
// writes quad 32w+q of word w). Order: corner rows ascending within a warp's run, warps in arrival order.
__global__ void __launch_bounds__(kThreads) k_compact_quads(MeshDims D, const uint32_t *__restrict__ bits, uint32_t *__restrict__ list,
                                                           uint32_t *__restrict__ count, unsigned long long *stamp) {
    pdl_trigger();
    pdl_wait();
    stage_stamp(stamp);
    const int lane = threadIdx.x & 31;
    const uint32_t nrows = (uint32_t)(D.ny + 1) * (uint32_t)(D.cz1 - D.cz0 + 1);
    const uint32_t nqw = (uint32_t)(D.nqx + 31) >> 5;  // 32-quad words per corner row
    const uint64_t nitems = (uint64_t)nrows * nqw;
    const uint64_t wpg = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t base = ((uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u; base < nitems; base += wpg * 32u) {
        const uint64_t item = base + lane;
        uint32_t needw = 0u, r = 0u, w = 0u;
        if (item < nitems) {
            r = (uint32_t)(item / nqw);
            w = (uint32_t)(item - (uint64_t)r * nqw);
            const int j = (int)(r % (uint32_t)(D.ny + 1));
            const int k = D.cz0 + (int)(r / (uint32_t)(D.ny + 1));
            const int by0 = j - 1 >= 0 ? (j - 1) >> 2 : -1, by1 = j < D.ny ? j >> 2 : -1;
            const int bz0 = k - 1 >= D.cz0 ? (k - 1) >> 2 : -1, bz1 = k < D.cz1 ? k >> 2 : -1;
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const int by = a ? by1 : by0;
                if (by < 0 || (a && by1 == by0)) continue;
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const int bz = c ? bz1 : bz0;
                    if (bz < 0 || (c && bz1 == bz0)) continue;
                    const uint32_t *row = bits + ((size_t)(bz - D.bz0) * D.nby + by) * D.nwx;
                    const uint32_t cur = (int)w < D.nwx ? row[w] : 0u;
                    const uint32_t prev = (w >= 1 && (int)(w - 1) < D.nwx) ? row[w - 1] : 0u;
                    needw |= cur | (cur << 1) | (prev >> 31);
                }
            }
            const int rem = D.nqx - 32 * (int)w;  // quads of this word that exist
            if (rem < 32) needw &= (1u << rem) - 1u;
        }
        const uint32_t pc = (uint32_t)__popc(needw);
        const uint32_t incl = warp_incl_scan(pc);
        const uint32_t wtot = __shfl_sync(0xffffffffu, incl, 31);
        if (wtot == 0u) continue;  // warp-uniform
        uint32_t wbase = 0u;
        if (lane == 0) wbase = atomicAdd(count, wtot);
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        uint32_t nonzero = __ballot_sync(0xffffffffu, needw != 0u);
        while (nonzero) {  // warp-uniform: only the words that hold survivors
            const int src = __ffs(nonzero) - 1;
            nonzero &= nonzero - 1u;
            const uint32_t word = __shfl_sync(0xffffffffu, needw, src);
            const uint32_t off = __shfl_sync(0xffffffffu, incl - pc, src);
            const uint32_t qid = __shfl_sync(0xffffffffu, r * (uint32_t)D.nqx + 32u * w, src);
            if ((word >> lane) & 1u) list[wbase + off + __popc(word & ((1u << lane) - 1u))] = qid + (uint32_t)lane;
        }
    }
}

// ---------------------------------------------------------------------------------------------- marching cubes
struct MCArgs {
    MeshDims D;
    float ox, oy, oz, res, cubeDiag;
    const float *grid;      // slab corner planes, plane 0 = corner plane cz0
    const uint32_t *mbits;  // prune bit rows (k_mask_bits); nullptr = FlatRenderer semantics (no prune)
    const uint8_t *t_ntri;  // marching-cubes tables in GLOBAL memory (divergent __constant__ reads serialise)
    const int8_t *t_tris;
    uint32_t *segcount;     // per 32-cell segment: triangle count (pass 1) / exclusive offset (pass 2)
    float *tris;            // 9 floats per triangle
    uint64_t tri_capacity;
    uint8_t *cases;         // optional nx*ny*(cz1-cz0) bytes (pass 1 only)
    uint32_t *overflow;     // set to 1 if a triangle did not fit
    uint32_t *seg_list;     // compact list of non-empty 32-cell segments (pass 1 -> pass 2)
    uint32_t *seg_count;
    uint8_t *seg_cases;     // optional: the 32 cube-case indices of seg_list[i] at [32*i, 32*i+32) (k_mc_count_tma -> k_mc_emit),
                            // so that pass 2 does not classify again
    unsigned long long *stamp;  // optional stage stamp slot of the kernel this struct is passed to
    // When the emit pass is the last kernel of the render its last CTA does k_finish_render's work (fin_ctr != nullptr):
    uint32_t *fin_ctr; volatile uint32_t *fin_hctr; int fin_nctr;
    unsigned long long *fin_scanstate; uint32_t fin_nstate;
    unsigned long long *fin_dstamp; volatile unsigned long long *fin_hstamp; int fin_nstamp;
    uint32_t *fin_done;         // CTAs that finished emitting (one of the fin_nctr counters: re-armed with them)
};

// The end of a render: publish the device counters (and stage stamps) into mapped pinned host memory with plain stores and
// re-arm counters + look-back scan state for the NEXT render. Run by all threads of ONE CTA.
__device__ __forceinline__ void finish_render_cta(uint32_t *__restrict__ d_ctr, volatile uint32_t *h_ctr, int nctr, unsigned long long *__restrict__ scanstate,
                                                  uint32_t nstate, unsigned long long *__restrict__ d_stamp, volatile unsigned long long *h_stamp, int nstamp) {
    if (d_stamp && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        d_stamp[nstamp - 1] = t;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < (uint32_t)nctr; i += blockDim.x) { h_ctr[i] = d_ctr[i]; d_ctr[i] = 0u; }
    for (uint32_t k = threadIdx.x; k < nstate; k += blockDim.x) scanstate[k] = 0ull;
    if (d_stamp && threadIdx.x < (uint32_t)nstamp) h_stamp[threadIdx.x] = d_stamp[threadIdx.x];
    // No system-scope fence here: the host (and the copy kernels of the multi-slab driver) read these words only behind the
    // stream's completion event, which orders them. With a fence the last CTA of a slab waited ~14 us for it whenever an earlier
    // slab's DMA was in flight (measured: second of two slabs seen at 221 instead of 236 us; GSDF_SYSFENCE builds it back in)
#ifdef GSDF_SYSFENCE
    __threadfence_system();
#endif
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

// Classification of one 32-cell segment of cell row (g00..g11 point at the four corner rows of the cube row).
// 4 coalesced row loads give every lane the x=cx column of its cube, the x=cx+1 column comes from the next lane by
// shuffle (corner order of flatrenderer.go:222-233); case index per marchcubes.go:39-44, reject rule
// flatrenderer.go:218-220. v[] = corner values 0..7.
__device__ __forceinline__ int mc_classify_segment(const float *g00, const float *g01, const float *g10, const float *g11, int cx, int nx,
                                                   bool act, float cubeDiag, float (&v)[8]) {
    const int lane = threadIdx.x & 31;
    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
    if (cx <= nx) { a00 = __ldg(g00 + cx); a01 = __ldg(g01 + cx); a10 = __ldg(g10 + cx); a11 = __ldg(g11 + cx); }
    float b00 = __shfl_down_sync(0xffffffffu, a00, 1), b01 = __shfl_down_sync(0xffffffffu, a01, 1);
    float b10 = __shfl_down_sync(0xffffffffu, a10, 1), b11 = __shfl_down_sync(0xffffffffu, a11, 1);
    if (lane == 31 && cx + 1 <= nx) { b00 = __ldg(g00 + cx + 1); b01 = __ldg(g01 + cx + 1); b10 = __ldg(g10 + cx + 1); b11 = __ldg(g11 + cx + 1); }
    v[0] = a00; v[1] = b00; v[2] = b01; v[3] = a01; v[4] = a10; v[5] = b10; v[6] = b11; v[7] = a11;
    int index = 0;
    if (act && !(fabsf(a00) > cubeDiag)) {
#pragma unroll
        for (int i = 0; i < 8; i++) index |= (v[i] < 0.f ? 1 : 0) << i;
    }
    return index;
}

// Pass 1: one warp per group of 4 consecutive 32-cell segments (= 32 prune blocks) of a cell row: one mask load +
// ballot decides which of its segments can hold surface; the others cost nothing. Writes the triangle count of every
// segment (the exclusive scan over this array gives output offsets in FlatRenderer cell order) and appends non-empty
// segments to a compact work list for pass 2.
__global__ void __launch_bounds__(kThreads) k_mc_count(MCArgs A) {
    __shared__ uint8_t s_ntri[256];
    pdl_trigger();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = A.t_ntri[i];
    pdl_wait();
    stage_stamp(A.stamp);
    __syncthreads();
    const MeshDims &D = A.D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nrows = (uint32_t)D.ny * (uint32_t)(D.cz1 - D.cz0);
    const uint32_t ngroups = (uint32_t)(D.nsx + 3) / 4;
    const uint32_t ntasks = nrows * ngroups;
    const uint32_t wpg = gridDim.x * (blockDim.x >> 5);
    const size_t sy = (size_t)D.pitch, sz = sy * (D.ny + 1);
    __shared__ uint32_t s_cnt[kThreads / 32];
    __shared__ uint32_t s_base;
    // CTA-uniform trip count so the 8 warps can share one atomicAdd per iteration for the work-list append
    for (uint32_t task0 = blockIdx.x * (blockDim.x >> 5); task0 < ntasks; task0 += wpg) {
        const uint32_t task = task0 + warp;
        uint32_t mine = 0u;  // lane sgm keeps the count of segment sgm
        uint32_t s0 = 0u;
        int nsg = 0;
        if (task < ntasks) {
            const uint32_t r = task / ngroups, g = task - r * ngroups;
            const int cy = (int)(r % (uint32_t)D.ny);
            const int czl = (int)(r / (uint32_t)D.ny);
            const int cz = D.cz0 + czl;
            const int b0 = (int)g * 32;
            uint32_t word = 0xffffffffu;
            if (A.mbits) word = A.mbits[((size_t)((cz >> 2) - D.bz0) * D.nby + (cy >> 2)) * D.nwx + g];
            s0 = r * (uint32_t)D.nsx + g * 4u;
            nsg = min(4, D.nsx - (int)g * 4);
            if (word == 0u) {
                if (lane < nsg) A.segcount[s0 + lane] = 0u;
                if (A.cases)
                    for (int c = lane; c < 128; c += 32) { const int cx = 4 * b0 + c; if (cx < D.nx) A.cases[(size_t)r * D.nx + cx] = 0; }
            } else {
                const float *g00 = A.grid + (size_t)czl * sz + (size_t)cy * sy;
                const float *g01 = g00 + sy, *g10 = g00 + sz, *g11 = g10 + sy;
#pragma unroll
                for (int sgm = 0; sgm < 4; sgm++) {
                    if (sgm >= nsg) break;
                    const int x0 = 4 * b0 + 32 * sgm;
                    const uint32_t bits = (word >> (8 * sgm)) & 0xffu;
                    const int cx = x0 + lane;
                    int index = 0;
                    uint32_t n = 0u;
                    if (bits != 0u) {
                        const bool act = cx < D.nx && ((bits >> (lane >> 2)) & 1u);
                        float v[8];
                        index = mc_classify_segment(g00, g01, g10, g11, cx, D.nx, act, A.cubeDiag, v);
                        n = s_ntri[index];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
                    }
                    if (A.cases && cx < D.nx) A.cases[(size_t)r * D.nx + cx] = (uint8_t)index;
                    if (lane == sgm) mine = n;
                }
                if (lane < nsg) A.segcount[s0 + lane] = mine;
            }
        }
        const unsigned nz = __ballot_sync(0xffffffffu, lane < nsg && mine != 0u);
        if (lane == 0) s_cnt[warp] = (uint32_t)__popc(nz);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < kThreads / 32; w++) { const uint32_t c = s_cnt[w]; s_cnt[w] = tot; tot += c; }
            s_base = tot ? atomicAdd(A.seg_count, tot) : 0u;
        }
        __syncthreads();
        if ((nz >> lane) & 1u) A.seg_list[s_base + s_cnt[warp] + __popc(nz & ((1u << lane) - 1u))] = s0 + lane;
        __syncthreads();
    }
}

// Pass 1, TMA form (default): one CTA per tile of 128 x 8 cells of one layer. The tile's 2 x 9 x 132 corner stencil is
// fetched by ONE 3-D tensor copy (cp.async.bulk.tensor.3d -> SASS UTMALDG) into shared memory, completion on an
// mbarrier; each of the 8 warps then classifies one cell row out of shared memory (8 conflict-free LDS per cell, no
// global loads, every lattice value read from L2 once per tile instead of up to four times). Tiles whose 2 x 32 prune
// blocks are all empty never issue the copy. Output is identical to k_mc_count.
constexpr int kTileX = 128, kTileY = 8, kBoxX = 132, kBoxY = 9, kBoxZ = 2;
__global__ void __launch_bounds__(256) k_mc_count_tma(const __grid_constant__ CUtensorMap tmap, MCArgs A) {
    __shared__ __align__(128) float s_tile[kBoxZ][kBoxY][kBoxX];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint8_t s_ntri[256];
    __shared__ uint32_t s_cnt[8];
    __shared__ uint32_t s_base;
    pdl_trigger();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = A.t_ntri[i];
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();  // the lattice (k_eval) and the prune bit rows are the predecessor's
    stage_stamp(A.stamp);
    __syncthreads();
    const MeshDims &D = A.D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ntx = (uint32_t)(D.nsx + 3) / 4, nty = (uint32_t)(D.ny + kTileY - 1) / kTileY, ntz = (uint32_t)(D.cz1 - D.cz0);
    const uint32_t ntiles = ntx * nty * ntz;
    uint32_t phase = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t tx = tile % ntx, ty = (tile / ntx) % nty, tz = tile / (ntx * nty);
        const int cz = D.cz0 + (int)tz, cy0 = (int)ty * kTileY;
        // prune words of the (up to) two block rows this tile touches
        uint32_t word0 = 0xffffffffu, word1 = 0xffffffffu;
        if (A.mbits) {
            const size_t rowb = ((size_t)((cz >> 2) - D.bz0) * D.nby + (cy0 >> 2)) * D.nwx + tx;
            word0 = A.mbits[rowb];
            word1 = ((cy0 >> 2) + 1 < D.nby) ? A.mbits[rowb + D.nwx] : 0u;
        }
        const bool active = (word0 | word1) != 0u;   // CTA-uniform
        if (active) {
            if (threadIdx.x == 0) {
                const uint32_t bytes = kBoxZ * kBoxY * kBoxX * 4;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                        smem_u32(&s_tile[0][0][0])),
                    "l"(&tmap), "r"((int)tx * kTileX), "r"(cy0), "r"((int)tz), "r"(bar)
                    : "memory");
            }
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "TW_LOOP:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra TW_DONE;\n\t"
                "bra TW_LOOP;\n\t"
                "TW_DONE:\n\t"
                "}" ::"r"(bar), "r"(phase)
                : "memory");
            phase ^= 1u;
        }
        // warp `warp` owns cell row cy of the tile
        const int cy = cy0 + warp;
        const uint32_t r = tz * (uint32_t)D.ny + (uint32_t)cy;
        const uint32_t s0 = r * (uint32_t)D.nsx + tx * 4u;
        const int nsg = cy < D.ny ? min(4, D.nsx - (int)tx * 4) : 0;
        const uint32_t word = ((cy >> 2) == (cy0 >> 2)) ? word0 : word1;
        uint32_t mine = 0u;
        uint32_t cases4 = 0u;  // this lane's cube-case index in each of the warp's four segments, one byte each
#pragma unroll
        for (int sgm = 0; sgm < 4; sgm++) {
            if (sgm >= nsg) break;
            const int lx = 32 * sgm + lane, cx = (int)tx * kTileX + lx;
            const uint32_t bits = (word >> (8 * sgm)) & 0xffu;
            int index = 0;
            uint32_t n = 0u;
            if (bits != 0u) {  // warp-uniform
                const bool act = cx < D.nx && ((bits >> (lane >> 2)) & 1u);
                const float v0 = s_tile[0][warp][lx], v1 = s_tile[0][warp][lx + 1], v2 = s_tile[0][warp + 1][lx + 1], v3 = s_tile[0][warp + 1][lx];
                const float v4 = s_tile[1][warp][lx], v5 = s_tile[1][warp][lx + 1], v6 = s_tile[1][warp + 1][lx + 1], v7 = s_tile[1][warp + 1][lx];
                if (act && !(fabsf(v0) > A.cubeDiag))
                    index = (v0 < 0.f ? 1 : 0) | (v1 < 0.f ? 2 : 0) | (v2 < 0.f ? 4 : 0) | (v3 < 0.f ? 8 : 0) | (v4 < 0.f ? 16 : 0) |
                            (v5 < 0.f ? 32 : 0) | (v6 < 0.f ? 64 : 0) | (v7 < 0.f ? 128 : 0);
                // most segments of a kept block hold no surface: one ballot decides, and the sum is one redux.sync
                if (__ballot_sync(0xffffffffu, index != 0 && index != 255) != 0u) n = __reduce_add_sync(0xffffffffu, (uint32_t)s_ntri[index]);
            }
            if (A.cases && cx < D.nx) A.cases[(size_t)r * D.nx + cx] = (uint8_t)index;
            if (lane == sgm) mine = n;
            cases4 |= (uint32_t)index << (8 * sgm);
        }
        if (lane < nsg) A.segcount[s0 + lane] = mine;
        const unsigned nz = __ballot_sync(0xffffffffu, lane < nsg && mine != 0u);
        if (lane == 0) s_cnt[warp] = (uint32_t)__popc(nz);
        __syncthreads();  // also: every warp is done reading s_tile before the next tile's copy may land
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; w++) { const uint32_t c = s_cnt[w]; s_cnt[w] = tot; tot += c; }
            s_base = tot ? atomicAdd(A.seg_count, tot) : 0u;
        }
        __syncthreads();
        const uint32_t lpos = s_base + s_cnt[warp];
        if ((nz >> lane) & 1u) A.seg_list[lpos + __popc(nz & ((1u << lane) - 1u))] = s0 + lane;
        if (A.seg_cases && nz) {  // warp-uniform: the case bytes of every non-empty segment, 32 coalesced bytes each
#pragma unroll
            for (int sgm = 0; sgm < 4; sgm++)
                if ((nz >> sgm) & 1u) A.seg_cases[(size_t)(lpos + __popc(nz & ((1u << sgm) - 1u))) * 32u + lane] = (uint8_t)(cases4 >> (8 * sgm));
        }
        __syncthreads();
    }
}

// Pass 1, TMA form, 4 cells per lane (default). Same tiles, same stencil copy and same outputs as k_mc_count_tma, but a
// warp classifies its whole 128-cell row in ONE pass: lane l owns the 4 cells of prune block l of the row (so bit l of
// the row's prune word is the lane's own verdict), reads its 2 x 2 x 5 corner values as four conflict-free 16-byte
// shared-memory loads plus the neighbour lane's first value (shuffle), turns each corner row into a 5-bit sign mask once
// and assembles the four cube-case indices from 2-bit slices of those masks (corner order flatrenderer.go:222-233,
// reject rule :218-220). 20 sign tests per 4 cells instead of 32, 4 loads instead of 32, one ballot instead of four;
// tiles without a kept block skip the copy, the barriers and the list append altogether.
__device__ __forceinline__ uint32_t rev2(uint32_t p) { return ((p & 1u) << 1) | (p >> 1); }
__global__ void __launch_bounds__(256, 8) k_mc_count_tma4(const __grid_constant__ CUtensorMap tmap, MCArgs A) {
    __shared__ __align__(128) float s_tile[kBoxZ][kBoxY][kBoxX];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint8_t s_ntri[256];
    __shared__ uint32_t s_cnt[8];
    __shared__ uint32_t s_base;
    pdl_trigger();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = A.t_ntri[i];
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();  // the lattice (k_eval) and the prune bit rows are the predecessor's
    stage_stamp(A.stamp);
    __syncthreads();
    const MeshDims &D = A.D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ntx = (uint32_t)(D.nsx + 3) / 4, nty = (uint32_t)(D.ny + kTileY - 1) / kTileY, ntz = (uint32_t)(D.cz1 - D.cz0);
    const uint32_t ntiles = ntx * nty * ntz;
    uint32_t phase = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t tx = tile % ntx, ty = (tile / ntx) % nty, tz = tile / (ntx * nty);
        const int cz = D.cz0 + (int)tz, cy0 = (int)ty * kTileY;
        uint32_t word0 = 0xffffffffu, word1 = 0xffffffffu;
        if (A.mbits) {
            const size_t rowb = ((size_t)((cz >> 2) - D.bz0) * D.nby + (cy0 >> 2)) * D.nwx + tx;
            word0 = A.mbits[rowb];
            word1 = ((cy0 >> 2) + 1 < D.nby) ? A.mbits[rowb + D.nwx] : 0u;
        }
        const int cy = cy0 + warp;                    // warp `warp` owns cell row cy of the tile
        const uint32_t r = tz * (uint32_t)D.ny + (uint32_t)cy;
        const uint32_t s0 = r * (uint32_t)D.nsx + tx * 4u;
        const int nsg = cy < D.ny ? min(4, D.nsx - (int)tx * 4) : 0;
        const int lx = 4 * lane, cx0 = (int)tx * kTileX + lx;
        if ((word0 | word1) == 0u) {                  // CTA-uniform: no kept block in the tile
            if (lane < nsg) A.segcount[s0 + lane] = 0u;
            if (A.cases && cy < D.ny) {
#pragma unroll
                for (int c = 0; c < 4; c++) if (cx0 + c < D.nx) A.cases[(size_t)r * D.nx + cx0 + c] = 0;
            }
            continue;
        }
        if (threadIdx.x == 0) {
            const uint32_t bytes = kBoxZ * kBoxY * kBoxX * 4;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                    smem_u32(&s_tile[0][0][0])),
                "l"(&tmap), "r"((int)tx * kTileX), "r"(cy0), "r"((int)tz), "r"(bar)
                : "memory");
        }
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "TW4_LOOP:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra TW4_DONE;\n\t"
            "bra TW4_LOOP;\n\t"
            "TW4_DONE:\n\t"
            "}" ::"r"(bar), "r"(phase)
            : "memory");
        phase ^= 1u;
        const uint32_t word = cy < D.ny ? (((cy >> 2) == (cy0 >> 2)) ? word0 : word1) : 0u;
        uint32_t cases4 = 0u;  // this lane's four cube-case indices, one byte each (cell cx0 + c in byte c)
        uint32_t mine = 0u;
        if (word != 0u) {      // warp-uniform
            uint32_t sm[2][2];
            float v0[4];
#pragma unroll
            for (int z = 0; z < 2; z++) {
#pragma unroll
                for (int y = 0; y < 2; y++) {
                    const float4 q = *reinterpret_cast<const float4 *>(&s_tile[z][warp + y][lx]);
                    float e = __shfl_down_sync(0xffffffffu, q.x, 1);
                    if (lane == 31) e = s_tile[z][warp + y][kTileX];
                    sm[z][y] = (q.x < 0.f ? 1u : 0u) | (q.y < 0.f ? 2u : 0u) | (q.z < 0.f ? 4u : 0u) | (q.w < 0.f ? 8u : 0u) | (e < 0.f ? 16u : 0u);
                    if (z == 0 && y == 0) { v0[0] = q.x; v0[1] = q.y; v0[2] = q.z; v0[3] = q.w; }
                }
            }
            const bool act = ((word >> lane) & 1u) != 0u;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                // corners 0..7 = (x c, y0, z0) (c+1, y0, z0) (c+1, y1, z0) (c, y1, z0) and the same on z1
                uint32_t idx = ((sm[0][0] >> c) & 3u) | (rev2((sm[0][1] >> c) & 3u) << 2) | (((sm[1][0] >> c) & 3u) << 4) | (rev2((sm[1][1] >> c) & 3u) << 6);
                if (!(act && cx0 + c < D.nx && !(fabsf(v0[c]) > A.cubeDiag))) idx = 0u;
                cases4 |= idx << (8 * c);
            }
            // some byte outside {0, 255}  <=>  (x ^ 0xff in every byte whose low bit is set) != 0
            const uint32_t mixed = cases4 ^ ((cases4 & 0x01010101u) * 255u);
            if (__ballot_sync(0xffffffffu, mixed != 0u) != 0u) {
                uint32_t n = (uint32_t)s_ntri[cases4 & 0xffu] + s_ntri[(cases4 >> 8) & 0xffu] + s_ntri[(cases4 >> 16) & 0xffu] + s_ntri[cases4 >> 24];
                n += __shfl_xor_sync(0xffffffffu, n, 1);
                n += __shfl_xor_sync(0xffffffffu, n, 2);
                n += __shfl_xor_sync(0xffffffffu, n, 4);   // every lane of an 8-lane group holds its segment's total
                mine = __shfl_sync(0xffffffffu, n, (lane & 3) * 8);  // lane sgm (< 4) keeps the count of segment sgm
            }
        }
        if (A.cases && cy < D.ny) {
#pragma unroll
            for (int c = 0; c < 4; c++) if (cx0 + c < D.nx) A.cases[(size_t)r * D.nx + cx0 + c] = (uint8_t)(cases4 >> (8 * c));
        }
        if (lane < nsg) A.segcount[s0 + lane] = mine;
        const unsigned nz = __ballot_sync(0xffffffffu, lane < nsg && mine != 0u);
        if (lane == 0) s_cnt[warp] = (uint32_t)__popc(nz);
        __syncthreads();  // also: every warp is done reading s_tile before the next tile's copy may land
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; w++) { const uint32_t c = s_cnt[w]; s_cnt[w] = tot; tot += c; }
            s_base = tot ? atomicAdd(A.seg_count, tot) : 0u;
        }
        __syncthreads();
        const uint32_t lpos = s_base + s_cnt[warp];
        if ((nz >> lane) & 1u) A.seg_list[lpos + __popc(nz & ((1u << lane) - 1u))] = s0 + lane;
        if (A.seg_cases) {  // the case bytes of every non-empty segment: lane l holds cells 4(l&7)..+3 of segment l>>3
            const int sgm = lane >> 3;
            if ((nz >> sgm) & 1u)
                *reinterpret_cast<uint32_t *>(A.seg_cases + (size_t)(lpos + __popc(nz & ((1u << sgm) - 1u))) * 32u + 4 * (lane & 7)) = cases4;
        }
        __syncthreads();
    }
}

// Pass 2: one warp per non-empty segment. segcount[] now holds exclusive triangle offsets. A surface usually crosses
// a row segment in only a few cells, so the segment's triangle VERTICES (3 per triangle) are dealt round-robin to
// the 32 lanes: each lane finds the owning cell of its vertex by a shuffle search over the inclusive scan of the
// per-cell triangle counts, interpolates that one edge and stores 3 floats.
__global__ void __launch_bounds__(kThreads) k_mc_emit(MCArgs A) {
    __shared__ uint8_t s_ntri[256];
    __shared__ __align__(16) int8_t s_tris[256 * 16];
    __shared__ float s_v[(kThreads / 32) * 8 * 32];  // [warp][corner][lane]
    pdl_trigger();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = A.t_ntri[i];
    for (int i = threadIdx.x; i < 256 * 16 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(s_tris)[i] = reinterpret_cast<const uint4 *>(A.t_tris)[i];
    pdl_wait();
    stage_stamp(A.stamp);
    __syncthreads();
    const MeshDims &D = A.D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nseg = *A.seg_count;
    const uint32_t wpg = gridDim.x * (blockDim.x >> 5);
    const size_t sy = (size_t)D.pitch, sz = sy * (D.ny + 1);
    float *vw = s_v + warp * 256;
    for (uint32_t t = blockIdx.x * (blockDim.x >> 5) + warp; t < nseg; t += wpg) {
        const uint32_t s = A.seg_list[t];
        const uint32_t r = s / (uint32_t)D.nsx;
        const int x0 = (int)(s - r * (uint32_t)D.nsx) << 5;
        const int cy = (int)(r % (uint32_t)D.ny);
        const int czl = (int)(r / (uint32_t)D.ny);
        const int cz = D.cz0 + czl;
        const float *g00 = A.grid + (size_t)czl * sz + (size_t)cy * sy;
        const float *g01 = g00 + sy, *g10 = g00 + sz, *g11 = g10 + sy;
        const int cx = x0 + lane;
        const bool precl = A.seg_cases != nullptr;  // launch-uniform: pass 1 left the case indices, corner values come from the lattice
        int index;
        if (precl) {
            index = (int)A.seg_cases[(size_t)t * 32u + lane];
        } else {
            bool act = cx < D.nx;
            if (A.mbits) act = act && bit_at(A.mbits + ((size_t)((cz >> 2) - D.bz0) * D.nby + (cy >> 2)) * D.nwx, min(cx, D.nx - 1) >> 2) != 0u;
            float v[8];
            index = mc_classify_segment(g00, g01, g10, g11, cx, D.nx, act, A.cubeDiag, v);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; i++) vw[i * 32 + lane] = v[i];
            __syncwarp();
        }
        const uint32_t n = s_ntri[index];
        const uint32_t incl = warp_incl_scan(n);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint64_t obase = (uint64_t)A.segcount[s];
        // corner positions, flatrenderer.go:235-247
        const float rr = A.res;
        const float py0 = A.oy + (float)cy * rr, pz0 = A.oz + (float)cz * rr;
        const float py1 = py0 + rr, pz1 = pz0 + rr;
        for (uint32_t ibase = 0; ibase < 3u * total; ibase += 32) {  // warp-uniform trip count: shuffles use all lanes
            const uint32_t item = ibase + lane;
            const bool live = item < 3u * total;
            const uint32_t tri = live ? item / 3u : total - 1u, j = live ? item - 3u * tri : 0u;
            // owner = first lane whose inclusive count exceeds tri (binary search by shuffle)
            int lo = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t probe = __shfl_sync(0xffffffffu, incl, lo + step - 1);
                if (probe <= tri) lo += step;
            }
            const int owner = lo;
            const uint32_t oincl = __shfl_sync(0xffffffffu, incl, owner);
            const uint32_t on = __shfl_sync(0xffffffffu, n, owner);
            const int oindex = __shfl_sync(0xffffffffu, index, owner);
            if (!live) continue;
            const uint32_t k = tri - (oincl - on);
            const uint64_t o = obase + tri;
            if (o >= A.tri_capacity) { *A.overflow = 1u; continue; }
            const float px0 = A.ox + (float)(x0 + owner) * rr, px1 = px0 + rr;
            // marchcubes.go:64-68: vertex j of the triangle is points[table[3k + 2 - j]]
            const int e = s_tris[16 * oindex + 3 * (int)k + 2 - (int)j];
            // edge -> corner pair (marchcubes.go:101-114), packed 4 bits per edge
            const int ca = (int)((0x321076543210ull >> (4 * e)) & 0xf), cb = (int)((0x765447650321ull >> (4 * e)) & 0xf);
            const float3 pa = make_float3((((ca + 1) >> 1) & 1) ? px1 : px0, ((ca >> 1) & 1) ? py1 : py0, (ca >> 2) ? pz1 : pz0);
            const float3 pb = make_float3((((cb + 1) >> 1) & 1) ? px1 : px0, ((cb >> 1) & 1) ? py1 : py0, (cb >> 2) ? pz1 : pz0);
            float va, vb;
            if (precl) {  // corner c of cell (x0+owner): x bit ((c+1)>>1)&1, y bit (c>>1)&1, z bit c>>2 (flatrenderer.go:222-233)
                va = __ldg(g00 + (size_t)(ca >> 2) * sz + (size_t)((ca >> 1) & 1) * sy + (x0 + owner + (((ca + 1) >> 1) & 1)));
                vb = __ldg(g00 + (size_t)(cb >> 2) * sz + (size_t)((cb >> 1) & 1) * sy + (x0 + owner + (((cb + 1) >> 1) & 1)));
            } else {
                va = vw[ca * 32 + owner]; vb = vw[cb * 32 + owner];
            }
            const float3 q = mc_interp(pa, pb, va, vb);
            float *dst = A.tris + 9 * o + 3 * j;
            dst[0] = q.x; dst[1] = q.y; dst[2] = q.z;
        }
    }
}

// Last node of a render: publishes the device counters into mapped pinned host memory with plain stores, then re-arms the
// state for the NEXT render (counters, look-back scan state) so that a render needs no separate clear pass.
// Why stores and not cudaMemcpyAsync / cudaMemsetAsync: small copies and memsets may be served by a copy engine, where they
// queue behind a device->host triangle read in flight (the 11 MB read of the previous Z-slab stalled the host's "how many
// triangles?" wait by 200 us, scripts/exp_pipe.py); a store from an SM does not.
__global__ void __launch_bounds__(256) k_finish_render(uint32_t *__restrict__ d_ctr, volatile uint32_t *h_ctr, int nctr,
                                                      unsigned long long *__restrict__ scanstate, uint32_t nstate,
                                                      unsigned long long *__restrict__ d_stamp, volatile unsigned long long *h_stamp, int nstamp) {
    pdl_wait();
    if (d_stamp) stage_stamp(d_stamp + nstamp - 1);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (uint32_t)nctr) { h_ctr[i] = d_ctr[i]; d_ctr[i] = 0u; }
    if (d_stamp && blockIdx.x == 0) {
        __syncthreads();
        if (threadIdx.x < (uint32_t)nstamp) h_stamp[threadIdx.x] = d_stamp[threadIdx.x];
    }
    for (uint32_t k = i; k < nstate; k += gridDim.x * blockDim.x) scanstate[k] = 0ull;
    __threadfence_system();
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

// Single-pass exclusive scan (decoupled look-back): tiles of 2048 items are claimed in order from an atomic ticket, each
// publishes its aggregate, looks back over its predecessors until it meets an inclusive prefix, publishes its own
// inclusive prefix and writes its items. State words carry an epoch so the array never needs clearing:
//   state = epoch << 34 | flag << 32 | value,  flag 1 = aggregate, 2 = inclusive prefix.
constexpr int kScanTile = kThreads * 8;
__global__ void __launch_bounds__(kThreads) k_scan_lookback(uint32_t *__restrict__ data, uint32_t n, unsigned long long *__restrict__ state,
                                                           uint32_t *__restrict__ ticket, uint32_t epoch, unsigned long long *__restrict__ total,
                                                           unsigned long long *stamp) {
    __shared__ uint32_t s_w[kThreads / 32];
    __shared__ uint32_t s_tile, s_prefix;
    pdl_trigger();
    pdl_wait();
    stage_stamp(stamp);
    const uint32_t ntiles = (n + kScanTile - 1) / kScanTile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= ntiles) return;
    const uint32_t base = tile * kScanTile + threadIdx.x * 8;
    uint32_t v[8], sum = 0;
    if (base + 8 <= n) {
        const uint4 a = *reinterpret_cast<const uint4 *>(data + base), b = *reinterpret_cast<const uint4 *>(data + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = base + i < n ? data[base + i] : 0u;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) sum += v[i];
    const uint32_t incl = warp_incl_scan(sum);
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t w = threadIdx.x < kThreads / 32 ? s_w[threadIdx.x] : 0u;
        const uint32_t wi = warp_incl_scan(w);
        if (threadIdx.x < kThreads / 32) s_w[threadIdx.x] = wi - w;
        const uint32_t agg = __shfl_sync(0xffffffffu, wi, 31);  // tile aggregate
        // lane 0 publishes, then the warp looks back 32 tiles at a time
        const unsigned long long tag = (unsigned long long)epoch << 34;
        if (threadIdx.x == 0) {
            const unsigned long long st = tag | ((tile == 0 ? 2ull : 1ull) << 32) | agg;
            atomicExch(&state[tile], st);
        }
        uint32_t prefix = 0;
        if (tile > 0) {
            int look = (int)tile - 1;
            for (;;) {
                const int idx = look - (int)threadIdx.x;
                unsigned long long st = 0;
                if (idx >= 0) {
                    do { st = *reinterpret_cast<volatile unsigned long long *>(&state[idx]); } while ((st >> 34) != epoch || ((st >> 32) & 3ull) == 0ull);
                }
                const uint32_t flag = idx >= 0 ? (uint32_t)((st >> 32) & 3ull) : 2u;  // before tile 0: inclusive prefix 0
                const uint32_t val = idx >= 0 ? (uint32_t)st : 0u;
                const unsigned incl_mask = __ballot_sync(0xffffffffu, flag == 2u);
                // take values up to and including the first inclusive prefix (lowest lane = nearest tile)
                const int stop = incl_mask ? __ffs(incl_mask) - 1 : 31;
                uint32_t part = (int)threadIdx.x <= stop ? val : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                prefix += part;
                if (incl_mask) break;
                look -= 32;
            }
            if (threadIdx.x == 0) atomicExch(&state[tile], tag | (2ull << 32) | (unsigned long long)(prefix + agg));
        }
        if (threadIdx.x == 0) {
            s_prefix = prefix;
            if (tile == ntiles - 1) *total = (unsigned long long)prefix + agg;
        }
    }
    __syncthreads();
    uint32_t run = s_prefix + s_w[threadIdx.x >> 5] + (incl - sum);
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { o[i] = run; run += v[i]; }
    if (base + 8 <= n) {
        *reinterpret_cast<uint4 *>(data + base) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4 *>(data + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) if (base + i < n) data[base + i] = o[i];
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
