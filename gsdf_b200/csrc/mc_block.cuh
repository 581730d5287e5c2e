// mc_block.cuh -- marching cubes over the KEPT 4x4x4-cell prune blocks only (default since round 2).
//
// The tile kernels (mc_kernels.cuh, mc_tile5.cuh) sweep whole rows of the lattice: on a pruned lattice a quarter of their
// lanes sit on kept blocks and a few per cent on cells that hold surface, and ncu shows both passes bound by instruction
// issue, not by memory. Here the unit of work is one kept block and nothing else is touched:
//   k_mesh_lists      (prune -> work lists) the quad list of the lattice evaluation AND the list of kept blocks
//   k_mc_blk_count    warp per kept block: its 5x5x5 corner stencil arrives in shared memory by ONE 3-D tensor copy
//                     (cp.async.bulk.tensor.3d, box 8x5x5, double-buffered per warp); 2 cells per lane are classified and
//                     the triangle count of each of the block's 16 cell rows goes to blkcnt[segment][slot] (one byte;
//                     a 32-cell row segment has 8 block slots). No atomics.
//   k_scan_seg        decoupled look-back scan whose input is computed on the fly: the count of a segment is the sum of the
//                     bytes of its KEPT slots (prune bit rows decide; stale bytes of other slots are never read as counts)
//   k_mc_blk_emit     warp per kept block again: rows without triangles cost one 16-byte load; the others re-classify
//                     their 64 cells from the staged stencil and deal the block's triangle vertices round-robin to the 32
//                     lanes (owner cell by a shuffle search over the inclusive scan of the per-cell counts -- the
//                     warp-level scan north_star asks for). A cell's triangles start at segoff[segment] + the kept slots
//                     in front of its block + the cells in front of it in its row: FlatRenderer cell order, deterministic.
// Output is bit-identical to k_mc_count / k_mc_emit.
#pragma once
#include "mc_kernels.cuh"

namespace gsdfk {

constexpr int kBlkBoxX = 8, kBlkBoxY = 5, kBlkBoxZ = 5;
constexpr uint32_t kBlkBoxBytes = kBlkBoxX * kBlkBoxY * kBlkBoxZ * 4;  // 800
constexpr uint32_t kBlkBufStride = 896;                                // 128-byte aligned per-warp stencil buffers
constexpr int kBlkWarps = 8;

struct BlkArgs {
    MeshDims D;
    float ox, oy, oz, res, cubeDiag;
    const uint32_t *mbits;     // prune bit rows, nullptr = every block kept (FlatRenderer)
    const uint8_t *childmask;  // plan ending with level 2: per block, which of its eight 2-cell cubes survived (nullptr: all)
    const uint32_t *blklist;   // kept blocks: (bzl * nby + by) * nbx + bx, bzl relative to D.bz0
    const uint32_t *nblk;      // device-side length of blklist
    uint2 *blkcnt;             // [segment] 8 bytes: triangles of block slot t in that 32-cell row segment
    unsigned long long *tilesum;  // != nullptr: the count pass also adds every row count to the sum of its scan tile (see scan_seg_tile)
    uint32_t *segoff;          // [segment] exclusive triangle offset (k_scan_seg)
    const uint8_t *t_ntri;
    const int8_t *t_tris;
    float *tris;
    uint64_t tri_capacity;
    uint8_t *cases;            // optional, zero-initialised by the host: nx*ny*(cz1-cz0) bytes
    uint32_t *overflow;
    unsigned long long *stamp;
    // the emit pass ends the render (see MCArgs)
    uint32_t *fin_ctr; volatile uint32_t *fin_hctr; int fin_nctr;
    unsigned long long *fin_scanstate; uint32_t fin_nstate;
    unsigned long long *fin_dstamp; volatile unsigned long long *fin_hstamp; int fin_nstamp;
    uint32_t *fin_done;
    // the emit pass runs the segment scan itself (no k_scan_seg launch): scan_state != nullptr
    unsigned long long *scan_state; uint32_t *scan_ticket, *scan_done; uint32_t scan_epoch, scan_nseg; unsigned long long *scan_total;};

// ---------------------------------------------------------------------------------------------- prune -> work lists
// Quad list exactly as k_compact_quads (lane per 32-quad word of a corner row), then the kept-block list (lane per 32-block
// word of a block row). bits == nullptr: no quad list, every block of the slab is listed (FlatRenderer).
// bits2 != nullptr (plan ending with level 2, PruneFine): a quad is needed iff a kept 2-CELL cube touches it -- quad m of
// corner row (j, k) is touched by the cubes 2m-1, 2m, 2m+1 of the level-2 rows (j-1)>>1, j>>1 x (k-1)>>1, k>>1.
// half != 0 (the consumer evaluates two corners per thread, GenGrid<2> with Lat::hq): the quad list holds HALF-quads,
// 2 * quad + h. A quad whose own block column is kept in one of the touching block rows lists both halves; a quad that is
// only needed as the last corner column of the kept block on its left lists its first half alone (its corners 4m+2, 4m+3
// -- and 4m+1 -- are read by nobody). `count` then counts half-quads.
__global__ void __launch_bounds__(kThreads) k_mesh_lists(MeshDims D, const uint32_t *__restrict__ bits, uint32_t *__restrict__ list,
                                                        uint32_t *__restrict__ count, uint32_t *__restrict__ blklist, uint32_t *__restrict__ nblk,
                                                        unsigned long long *stamp, const uint32_t *__restrict__ bits2, int half) {
    pdl_trigger();
    pdl_wait();
    stage_stamp(stamp);
    const int lane = threadIdx.x & 31;
    const uint64_t wpg = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t w0 = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // one index space for both lists: [0, nq_pad) the 32-quad words (padded to whole warps), then the block words -- a warp
    // builds one list or the other, so the two list-building chains (loads -> scan -> atomic -> stores) run side by side
    const uint32_t nrows = (uint32_t)(D.ny + 1) * (uint32_t)(D.cz1 - D.cz0 + 1);
    const uint32_t nqw = (uint32_t)(D.nqx + 31) >> 5;  // 32-quad words per corner row
    const uint64_t nq_items = (bits && list) ? (uint64_t)nrows * nqw : 0ull;
    const uint64_t nq_pad = (nq_items + 31ull) & ~31ull;
    const uint64_t nb_items = (uint64_t)D.nbz * D.nby * D.nwx;
    for (uint64_t base = w0 * 32u; base < nq_pad + nb_items; base += wpg * 32u) {
        if (base < nq_pad) {
            const uint64_t nitems = nq_items;
            const uint64_t item = base + lane;
            uint32_t needw = 0u, fullw = 0u, r = 0u, w = 0u;
            if (item < nitems) {
                r = (uint32_t)(item / nqw);
                w = (uint32_t)(item - (uint64_t)r * nqw);
                const int j = (int)(r % (uint32_t)(D.ny + 1));
                const int k = D.cz0 + (int)(r / (uint32_t)(D.ny + 1));
                const int sh = bits2 ? 1 : 2;  // log2 of the cube width the rows describe
                const int by0 = j - 1 >= 0 ? (j - 1) >> sh : -1, by1 = j < D.ny ? j >> sh : -1;
                const int bz0 = k - 1 >= D.cz0 ? (k - 1) >> sh : -1, bz1 = k < D.cz1 ? k >> sh : -1;
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    const int by = a ? by1 : by0;
                    if (by < 0 || (a && by1 == by0)) continue;
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const int bz = c ? bz1 : bz0;
                        if (bz < 0 || (c && bz1 == bz0)) continue;
                        if (bits2) {
                            const uint32_t *row = bits2 + ((size_t)(bz - 2 * D.bz0) * (2 * D.nby) + by) * (2 * D.nwx);
                            const uint32_t e = (int)w < D.nwx ? row[2 * w] : 0u, o = (int)w < D.nwx ? row[2 * w + 1] : 0u;
                            const uint32_t oprev = (w >= 1 && (int)(w - 1) < D.nwx) ? row[2 * (w - 1) + 1] : 0u;
                            needw |= e | o | (o << 1) | (oprev >> 31);
                            fullw = needw;  // (2-cell cubes: every listed quad keeps both halves)
                        } else {
                            const uint32_t *row = bits + ((size_t)(bz - D.bz0) * D.nby + by) * D.nwx;
                            const uint32_t cur = (int)w < D.nwx ? row[w] : 0u;
                            const uint32_t prev = (w >= 1 && (int)(w - 1) < D.nwx) ? row[w - 1] : 0u;
                            needw |= cur | (cur << 1) | (prev >> 31);
                            fullw |= cur;
                        }
                    }
                }
                const int rem = D.nqx - 32 * (int)w;  // quads of this word that exist
                if (rem < 32) needw &= (1u << rem) - 1u;
                fullw = half ? (fullw & needw) : 0u;
            }
            // list entries of this lane's word: one per needed quad, two where both halves are listed
            const uint32_t pc = (uint32_t)__popc(needw) + (uint32_t)__popc(fullw);
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
                if (!half) {
                    if ((word >> lane) & 1u) list[wbase + off + __popc(word & ((1u << lane) - 1u))] = qid + (uint32_t)lane;
                } else {
                    const uint32_t fw = __shfl_sync(0xffffffffu, fullw, src);
                    if ((word >> lane) & 1u) {
                        const uint32_t below = (1u << lane) - 1u;
                        const uint32_t at = wbase + off + (uint32_t)__popc(word & below) + (uint32_t)__popc(fw & below);
                        const uint32_t hq = 2u * (qid + (uint32_t)lane);
                        list[at] = hq;
                        if ((fw >> lane) & 1u) list[at + 1u] = hq + 1u;
                    }
                }
            }
        } else {  // kept blocks
            const uint64_t nitems = nb_items;
            const uint64_t item = base - nq_pad + lane;
            uint32_t word = 0u, brow = 0u, w = 0u;
            if (item < nitems) {
                brow = (uint32_t)(item / (uint32_t)D.nwx);
                w = (uint32_t)(item - (uint64_t)brow * (uint32_t)D.nwx);
                word = bits ? bits[item] : 0xffffffffu;
                const int rem = D.nbx - 32 * (int)w;
                if (rem < 32) word &= (1u << rem) - 1u;
            }
            const uint32_t pc = (uint32_t)__popc(word);
            const uint32_t incl = warp_incl_scan(pc);
            const uint32_t wtot = __shfl_sync(0xffffffffu, incl, 31);
            if (wtot == 0u) continue;
            uint32_t wbase = 0u;
            if (lane == 0) wbase = atomicAdd(nblk, wtot);
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            uint32_t nonzero = __ballot_sync(0xffffffffu, word != 0u);
            while (nonzero) {
                const int src = __ffs(nonzero) - 1;
                nonzero &= nonzero - 1u;
                const uint32_t wd = __shfl_sync(0xffffffffu, word, src);
                const uint32_t off = __shfl_sync(0xffffffffu, incl - pc, src);
                const uint32_t bid = __shfl_sync(0xffffffffu, brow * (uint32_t)D.nbx + 32u * w, src);
                if ((wd >> lane) & 1u) blklist[wbase + off + __popc(wd & ((1u << lane) - 1u))] = bid + (uint32_t)lane;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- per-block helpers
struct BlkPos {
    int bx, by, bzl;   // block coordinates (bzl relative to D.bz0)
    int x0, y0, z0;    // first cell (global cell coordinates)
};
__device__ __forceinline__ BlkPos blk_decode(const MeshDims &D, uint32_t bid) {
    BlkPos b;
    const uint32_t t = fastdiv(bid, D.nbx_mul, D.nbx_shr);
    b.bx = (int)(bid - t * (uint32_t)D.nbx);
    b.bzl = (int)fastdiv(t, D.nby_mul, D.nby_shr);
    b.by = (int)(t - (uint32_t)b.bzl * (uint32_t)D.nby);
    b.x0 = 4 * b.bx; b.y0 = 4 * b.by; b.z0 = 4 * (D.bz0 + b.bzl);
    return b;
}
// stencil copy of block b into `buf` (corner (x0+i, y0+j, z0+k) lands at buf[(k*5 + j)*8 + i]); plane index relative to cz0
__device__ __forceinline__ void blk_issue(const CUtensorMap *tmap, float *buf, uint32_t bar, const MeshDims &D, const BlkPos &b) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kBlkBoxBytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(buf)),
                 "l"(tmap), "r"(b.x0), "r"(b.y0), "r"(b.z0 - D.cz0), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void blk_wait(uint32_t bar, uint32_t phase) {
    if ((threadIdx.x & 31) == 0) asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "BW_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra BW_DONE;\n\t"
        "bra BW_LOOP;\n\t"
        "BW_DONE:\n\t"
        "}" ::"r"(bar), "r"(phase)
        : "memory");
    __syncwarp();
}
// Cube-case index of cell (lx, ly, lz) of the block from the staged stencil; corner order flatrenderer.go:222-233, reject
// rule :218-220, case index marchcubes.go:39-44. `valid`: the cell exists in the slab.
__device__ __forceinline__ int blk_case(const float *buf, int lx, int ly, int lz, bool valid, float cubeDiag) {
    const float *c = buf + (lz * 5 + ly) * 8 + lx;
    const float v0 = c[0], v1 = c[1], v2 = c[9], v3 = c[8], v4 = c[40], v5 = c[41], v6 = c[49], v7 = c[48];
    if (!valid || fabsf(v0) > cubeDiag) return 0;
    return (v0 < 0.f ? 1 : 0) | (v1 < 0.f ? 2 : 0) | (v2 < 0.f ? 4 : 0) | (v3 < 0.f ? 8 : 0) | (v4 < 0.f ? 16 : 0) | (v5 < 0.f ? 32 : 0) |
           (v6 < 0.f ? 64 : 0) | (v7 < 0.f ? 128 : 0);
}
// 32-cell row segment and block slot of cell row (cy, cz) of block column bx
__device__ __forceinline__ uint32_t blk_segment(const MeshDims &D, int bx, int cy, int cz) {
    return ((uint32_t)(cz - D.cz0) * (uint32_t)D.ny + (uint32_t)cy) * (uint32_t)D.nsx + (uint32_t)(bx >> 3);
}

// ---------------------------------------------------------------------------------------------- pass 1
// Pass 1 of one block from its staged stencil. Lane l: cell column lx = l & 3, row ly = (l >> 2) & 3, layers lz = l >> 4 and
// lz + 2. Writes the triangle count of each of the block's 16 cell rows (and the case indices in parity mode).
__device__ __forceinline__ void blk_count_block(const BlkArgs &A, const BlkPos &b, uint32_t cm, const float *buf, const uint8_t *s_ntri, int lane) {
    const MeshDims &D = A.D;
    const int lx = lane & 3, ly = (lane >> 2) & 3, lz = lane >> 4;
    const int cx = b.x0 + lx, cy = b.y0 + ly;
    const bool okxy = cx < D.nx && cy < D.ny;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int lzz = lz + 2 * h, cz = b.z0 + lzz;
        // (h = child layer dz: lz is 0 or 1) corners of a dropped 2-cell cube were never evaluated: its cells hold no surface
        const bool valid = okxy && cz >= D.cz0 && cz < D.cz1 && ((cm >> ((lx >> 1) + 2 * (ly >> 1) + 4 * h)) & 1u);
        const int idx = blk_case(buf, lx, ly, lzz, valid, A.cubeDiag);
        uint32_t n = s_ntri[idx];
        n += __shfl_xor_sync(0xffffffffu, n, 1);
        n += __shfl_xor_sync(0xffffffffu, n, 2);  // the four cells of the row
        if (cy < D.ny && cz >= D.cz0 && cz < D.cz1) {
            if (lx == 0) {
                const uint32_t seg = blk_segment(D, b.bx, cy, cz);
                reinterpret_cast<uint8_t *>(A.blkcnt + seg)[b.bx & 7] = (uint8_t)n;
                if (A.tilesum && n) atomicAdd(A.tilesum + seg / kScanTile, (unsigned long long)n);  // (no return value: a reduction at L2)
            }
            if (A.cases && cx < D.nx) A.cases[((size_t)(cz - D.cz0) * D.ny + cy) * D.nx + cx] = (uint8_t)idx;
        }
    }
}

__global__ void __launch_bounds__(kBlkWarps * 32) k_mc_blk_count(const __grid_constant__ CUtensorMap tmap, BlkArgs A) {
    __shared__ __align__(128) uint8_t s_buf[kBlkWarps][2][kBlkBufStride];
    __shared__ __align__(8) uint64_t s_bar[kBlkWarps][2];
    __shared__ uint8_t s_ntri[256];
    pdl_trigger();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = A.t_ntri[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t bar0 = smem_u32(&s_bar[warp][0]), bar1 = smem_u32(&s_bar[warp][1]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();  // lattice, bit rows and block list are the predecessors'
    stage_stamp(A.stamp);
    __syncthreads();
    const MeshDims &D = A.D;
    const uint32_t nblk = *A.nblk;
    const uint32_t stride = gridDim.x * kBlkWarps;
    uint32_t phases = 0u;  // bit b = parity to wait for on buffer b
    uint32_t it = blockIdx.x * kBlkWarps + warp;
    int cur = 0;
    if (it < nblk && lane == 0) blk_issue(&tmap, reinterpret_cast<float *>(s_buf[warp][0]), bar0, D, blk_decode(D, A.blklist[it]));
    for (; it < nblk; it += stride, cur ^= 1) {
        const BlkPos b = blk_decode(D, A.blklist[it]);
        if (it + stride < nblk && lane == 0)  // the other buffer was released by the __syncwarp that ended the previous iteration
            blk_issue(&tmap, reinterpret_cast<float *>(s_buf[warp][cur ^ 1]), cur ? bar0 : bar1, D, blk_decode(D, A.blklist[it + stride]));
        const uint32_t cm = A.childmask ? A.childmask[A.blklist[it]] : 0xffu;
        blk_wait(cur ? bar1 : bar0, (phases >> cur) & 1u);
        phases ^= 1u << cur;
        blk_count_block(A, b, cm, reinterpret_cast<const float *>(s_buf[warp][cur]), s_ntri, lane);
        __syncwarp();  // everybody is done with buf before it is refilled two iterations later
    }
}

// ---------------------------------------------------------------------------------------------- scan over segments
// Triangle count of a segment = sum of the bytes of its kept block slots. kb = kept bits of the segment's 8 slots.
__device__ __forceinline__ uint32_t seg_kept_bits_at(const MeshDims &D, const uint32_t *mbits, int cy, int cz, uint32_t sx) {
    uint32_t kb = 0xffu;
    if (mbits) kb = (mbits[((size_t)((cz >> 2) - D.bz0) * D.nby + (cy >> 2)) * D.nwx + (sx >> 2)] >> (8u * (sx & 3u))) & 0xffu;
    const int rem = D.nbx - 8 * (int)sx;  // block slots of this segment that exist
    if (rem < 8) kb &= (1u << rem) - 1u;
    return kb;
}
__device__ __forceinline__ uint32_t seg_kept_bits(const MeshDims &D, const uint32_t *mbits, uint32_t row, uint32_t sx) {
    const int cy = (int)(row % (uint32_t)D.ny), cz = D.cz0 + (int)(row / (uint32_t)D.ny);
    uint32_t kb = 0xffu;
    if (mbits) kb = (mbits[((size_t)((cz >> 2) - D.bz0) * D.nby + (cy >> 2)) * D.nwx + (sx >> 2)] >> (8u * (sx & 3u))) & 0xffu;
    const int rem = D.nbx - 8 * (int)sx;  // block slots of this segment that exist
    if (rem < 8) kb &= (1u << rem) - 1u;
    return kb;
}
// sum of the count bytes of the kept slots t < upto: the kept bits are spread to one byte each (bit i of a nibble -> byte i) and
// dotted with the eight count bytes, two dp4a instead of an 8-step select-and-add chain
__device__ __forceinline__ uint32_t spread4(uint32_t nibble) { return ((nibble & 0xfu) * 0x00204081u) & 0x01010101u; }
__device__ __forceinline__ uint32_t seg_masked_sum(uint2 c, uint32_t kb, int upto = 8) {
    const uint32_t m = kb & ((1u << upto) - 1u);
    return __dp4a(c.x, spread4(m), 0u) + __dp4a(c.y, spread4(m >> 4), 0u);
}

// One tile (kScanTile segments) of the segment scan, by all kThreads threads of a CTA. Two ways to the tile's prefix:
//   tilesums: the count pass has already added every row count to state[tile of its segment] (a plain sum, no flags), so the
//             prefix is the sum of the words in front -- one coalesced read, nothing to wait for (used up to kTileSumMax tiles);
//   else:     decoupled look-back; flag values of a tile's state word: 1 = aggregate published, 2 = inclusive prefix published.
// done_counter (scan inside the emit pass): += 1 once the tile's offsets are in segoff.
constexpr uint32_t kTileSumMax = 1024;
__device__ __forceinline__ void scan_seg_tile(const MeshDims &D, const uint32_t *__restrict__ mbits, const uint2 *__restrict__ blkcnt, uint32_t *__restrict__ segoff,
                                              uint32_t n, unsigned long long *__restrict__ state, uint32_t epoch, unsigned long long *__restrict__ total,
                                              uint32_t tile, uint32_t *s_w, uint32_t *s_prefix, uint32_t *done_counter, bool tilesums) {
    const uint32_t ntiles = (n + kScanTile - 1) / kScanTile;
    const uint32_t base = tile * kScanTile + threadIdx.x * 8;
    uint32_t v[8], sum = 0;
    {
        uint32_t row = base / (uint32_t)D.nsx, sx = base - row * (uint32_t)D.nsx;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            v[i] = 0u;
            if (base + i < n) {
                const uint32_t kb = seg_kept_bits(D, mbits, row, sx);
                if (kb) v[i] = seg_masked_sum(__ldcg(blkcnt + base + i), kb);
            }
            if (++sx == (uint32_t)D.nsx) { sx = 0u; row++; }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) sum += v[i];
    const uint32_t incl = warp_incl_scan(sum);
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (tilesums) {
        uint32_t part = 0;
        for (uint32_t t = threadIdx.x; t < tile; t += kThreads) part += (uint32_t)__ldcg(state + t);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        __shared__ uint32_t s_part[kThreads / 32];
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t w = threadIdx.x < kThreads / 32 ? s_w[threadIdx.x] : 0u;
            const uint32_t wi = warp_incl_scan(w);
            if (threadIdx.x < kThreads / 32) s_w[threadIdx.x] = wi - w;
            const uint32_t agg = __shfl_sync(0xffffffffu, wi, 31);  // tile aggregate
            uint32_t prefix = threadIdx.x < kThreads / 32 ? s_part[threadIdx.x] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) prefix += __shfl_xor_sync(0xffffffffu, prefix, o);
            if (threadIdx.x == 0) {
                *s_prefix = prefix;
                if (tile == ntiles - 1) *total = (unsigned long long)prefix + agg;
            }
        }
    } else {
    const unsigned long long tag = (unsigned long long)epoch << 34;
    if (threadIdx.x < 32) {
        const uint32_t w = threadIdx.x < kThreads / 32 ? s_w[threadIdx.x] : 0u;
        const uint32_t wi = warp_incl_scan(w);
        if (threadIdx.x < kThreads / 32) s_w[threadIdx.x] = wi - w;
        const uint32_t agg = __shfl_sync(0xffffffffu, wi, 31);  // tile aggregate
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
            *s_prefix = prefix;
            if (tile == ntiles - 1) *total = (unsigned long long)prefix + agg;
        }
    }
    }
    __syncthreads();
    uint32_t run = *s_prefix + s_w[threadIdx.x >> 5] + (incl - sum);
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { o[i] = run; run += v[i]; }
    if (base + 8 <= n) {
        *reinterpret_cast<uint4 *>(segoff + base) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4 *>(segoff + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) if (base + i < n) segoff[base + i] = o[i];
    }
    if (done_counter) {  // release the tile's offsets to the emit phase of every CTA
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(done_counter, 1u);
    }
}

__global__ void __launch_bounds__(kThreads) k_scan_seg(MeshDims D, const uint32_t *__restrict__ mbits, const uint2 *__restrict__ blkcnt, uint32_t *__restrict__ segoff,
                                                      uint32_t n, unsigned long long *__restrict__ state, uint32_t *__restrict__ ticket, uint32_t epoch,
                                                      unsigned long long *__restrict__ total, unsigned long long *stamp, int tilesums) {
    __shared__ uint32_t s_w[kThreads / 32];
    __shared__ uint32_t s_tile, s_prefix[2];
    pdl_trigger();
    pdl_wait();
    stage_stamp(stamp);
    const uint32_t ntiles = (n + kScanTile - 1) / kScanTile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= ntiles) return;
    scan_seg_tile(D, mbits, blkcnt, segoff, n, state, epoch, total, tile, s_w, s_prefix, nullptr, tilesums != 0);
}

// ---------------------------------------------------------------------------------------------- pass 2
constexpr int kBlkListCap = 32 * 5;  // triangles of one half-block (32 cells, at most 5 each)

// What pass 1 left for the lane's two cell rows (ly, lz) and (ly, lz + 2) of block b: the row's triangle count and where
// the row's triangles start in the output (segment offset + the kept block slots in front of this one). cg: the offsets were
// written during this kernel (scan inside the pass): read them around L1. (Loading the rows of block i+1 before block i is
// emitted was measured: 63 instead of 40 registers, no gain.)
__device__ __forceinline__ void blk_emit_rows(const BlkArgs &A, const BlkPos &b, int lane, bool cg, uint32_t (&rown)[2], uint32_t (&rowbase)[2]) {
    const MeshDims &D = A.D;
    // a block has 16 cell rows and every lane needs two of them, shared with the three other lanes of its row: lane l looks up
    // row l & 15 (ry = l & 3, rz = (l >> 2) & 3) once and the lanes pick their rows up by shuffle
    const int ry = lane & 3, rz = (lane >> 2) & 3;
    const int cy = b.y0 + ry, cz = b.z0 + rz;
    uint32_t rn = 0u, rb = 0u;
    if (cy < D.ny && cz >= D.cz0 && cz < D.cz1) {
        const uint32_t seg = blk_segment(D, b.bx, cy, cz);
        const uint2 c = cg ? __ldcg(A.blkcnt + seg) : A.blkcnt[seg];
        const int slot = b.bx & 7;
        rn = ((slot < 4 ? c.x : c.y) >> (8 * (slot & 3))) & 0xffu;
        if (rn) {  // triangles of the kept slots in front of this block, behind the segment's offset
            const uint32_t kb = seg_kept_bits_at(D, A.mbits, cy, cz, (uint32_t)(b.bx >> 3));
            const uint32_t so = cg ? __ldcg(A.segoff + seg) : A.segoff[seg];
            rb = so + seg_masked_sum(c, kb, slot);
        }
    }
    const int ly = (lane >> 2) & 3, lz = lane >> 4;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int src = ly + 4 * (lz + 2 * h);
        rown[h] = __shfl_sync(0xffffffffu, rn, src);
        rowbase[h] = __shfl_sync(0xffffffffu, rb, src);
    }
}

// The triangles of block b from its staged stencil (warp-uniform call). Per half-block (h = 0: layers 0, 1; h = 1: layers
// 2, 3) every lane lists the triangles of its cell at the positions the warp scan of the counts gives (the warp-level scan
// north_star asks for) -- one entry per triangle: output slot, cube case, triangle number within the cell, cell -- and the
// vertices are then dealt round-robin, lane = vertex: a vertex finds its triangle with one shared-memory load. A cell's
// triangles start at its row's base + the cells in front of it in the row: FlatRenderer cell order, deterministic.
__device__ __forceinline__ void blk_emit_tris(const BlkArgs &A, const BlkPos &b, uint32_t cm, const float *buf, const uint8_t *s_ntri, const int8_t *s_tris,
                                              uint2 *lst, int lane, const uint32_t (&rown)[2], const uint32_t (&rowbase)[2]) {
    const MeshDims &D = A.D;
    const int lx = lane & 3, ly = (lane >> 2) & 3, lz = lane >> 4;
    const int cx = b.x0 + lx, cy = b.y0 + ly;
    const bool okxy = cx < D.nx && cy < D.ny;
    const float rr = A.res;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (__ballot_sync(0xffffffffu, rown[h] != 0u) == 0u) continue;  // warp-uniform
        const int lzz = lz + 2 * h, cz = b.z0 + lzz;
        const bool valid = okxy && cz >= D.cz0 && cz < D.cz1 && rown[h] != 0u && ((cm >> ((lx >> 1) + 2 * (ly >> 1) + 4 * h)) & 1u);
        const int index = valid ? blk_case(buf, lx, ly, lzz, true, A.cubeDiag) : 0;
        const uint32_t n = s_ntri[index];
        // cells in front of this one in its row: exclusive prefix over the 4 lanes of the row
        const uint32_t a1 = __shfl_up_sync(0xffffffffu, n, 1), a2 = __shfl_up_sync(0xffffffffu, n, 2), a3 = __shfl_up_sync(0xffffffffu, n, 3);
        const uint32_t obase = rowbase[h] + (lx >= 1 ? a1 : 0u) + (lx >= 2 ? a2 : 0u) + (lx >= 3 ? a3 : 0u);
        const uint32_t incl = warp_incl_scan(n);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        for (uint32_t k = 0; k < n; k++) lst[incl - n + k] = make_uint2(obase + k, (uint32_t)index | (k << 8) | ((uint32_t)lane << 11));
        __syncwarp();
        for (uint32_t item = lane; item < 3u * total; item += 32) {
            const uint32_t tri = item / 3u, j = item - 3u * tri;
            const uint2 ent = lst[tri];
            const uint64_t o = ent.x;
            if (o >= A.tri_capacity) { *A.overflow = 1u; continue; }
            const int oindex = (int)(ent.y & 0xffu);
            const uint32_t kk = (ent.y >> 8) & 7u;
            const int owner = (int)(ent.y >> 11);
            const int olx = owner & 3, oly = (owner >> 2) & 3, olz = (owner >> 4) + 2 * h;
            // corner positions, flatrenderer.go:235-247
            const float px0 = A.ox + (float)(b.x0 + olx) * rr, px1 = px0 + rr;
            const float py0 = A.oy + (float)(b.y0 + oly) * rr, py1 = py0 + rr;
            const float pz0 = A.oz + (float)(b.z0 + olz) * rr, pz1 = pz0 + rr;
            // marchcubes.go:64-68: vertex j of the triangle is points[table[3k + 2 - j]]
            const int e = s_tris[16 * oindex + 3 * (int)kk + 2 - (int)j];
            // edge -> corner pair (marchcubes.go:101-114), packed 4 bits per edge
            const int ca = (int)((0x321076543210ull >> (4 * e)) & 0xf), cb = (int)((0x765447650321ull >> (4 * e)) & 0xf);
            const float3 pa = make_float3((((ca + 1) >> 1) & 1) ? px1 : px0, ((ca >> 1) & 1) ? py1 : py0, (ca >> 2) ? pz1 : pz0);
            const float3 pb = make_float3((((cb + 1) >> 1) & 1) ? px1 : px0, ((cb >> 1) & 1) ? py1 : py0, (cb >> 2) ? pz1 : pz0);
            // corner c of the cell: x bit ((c+1)>>1)&1, y bit (c>>1)&1, z bit c>>2 (flatrenderer.go:222-233)
            const float *cc = buf + (olz * 5 + oly) * 8 + olx;
            const float va = cc[(ca >> 2) * 40 + ((ca >> 1) & 1) * 8 + (((ca + 1) >> 1) & 1)];
            const float vb = cc[(cb >> 2) * 40 + ((cb >> 1) & 1) * 8 + (((cb + 1) >> 1) & 1)];
            const float3 q = mc_interp(pa, pb, va, vb);
            float *dst = A.tris + 9 * o + 3 * j;
            dst[0] = q.x; dst[1] = q.y; dst[2] = q.z;
        }
        __syncwarp();  // the list is rewritten by the next half-block
    }
}

// The scan phase of a pass that scans by itself: CTAs take scan tiles by ticket until none is left. A tile only ever waits
// for tiles with smaller tickets, i.e. for CTAs that are already running, so no co-residency of the whole grid is needed;
// then one thread per CTA waits (with back-off) until every tile has written its offsets. All threads of the CTA call.
__device__ __forceinline__ void blk_scan_phase(const BlkArgs &A, uint32_t *s_w, uint32_t *s_tile, uint32_t *s_prefix) {
    const uint32_t ntiles = (A.scan_nseg + kScanTile - 1) / kScanTile;
    for (;;) {
        if (threadIdx.x == 0) *s_tile = *reinterpret_cast<volatile uint32_t *>(A.scan_ticket) >= ntiles ? ntiles : atomicAdd(A.scan_ticket, 1u);
        __syncthreads();
        const uint32_t tile = *s_tile;
        if (tile >= ntiles) break;
        scan_seg_tile(A.D, A.mbits, A.blkcnt, A.segoff, A.scan_nseg, A.scan_state, A.scan_epoch, A.scan_total, tile, s_w, s_prefix, A.scan_done, A.tilesum != nullptr);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t d;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(d) : "l"(A.scan_done) : "memory");
            if (d >= ntiles) break;
            __nanosleep(64);
        }
    }
    __syncthreads();
}

// The last CTA of the render's last kernel publishes the counters and re-arms the state (k_finish_render's work, without
// its launch). All threads of the CTA call.
__device__ __forceinline__ void blk_finish(const BlkArgs &A, uint32_t *s_last) {
    if (!A.fin_ctr) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        *s_last = atomicAdd(A.fin_done, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (*s_last) {
        __threadfence();
        finish_render_cta(A.fin_ctr, A.fin_hctr, A.fin_nctr, A.fin_scanstate, A.fin_nstate, A.fin_dstamp, A.fin_hstamp, A.fin_nstamp);
    }
}

__global__ void __launch_bounds__(kBlkWarps * 32) k_mc_blk_emit(const __grid_constant__ CUtensorMap tmap, BlkArgs A) {
    __shared__ __align__(128) uint8_t s_buf[kBlkWarps][2][kBlkBufStride];
    __shared__ __align__(8) uint64_t s_bar[kBlkWarps][2];
    __shared__ uint8_t s_ntri[256];
    __shared__ __align__(16) int8_t s_tris[256 * 16];
    __shared__ uint2 s_list[kBlkWarps][kBlkListCap];
    __shared__ uint32_t s_w[kThreads / 32];
    __shared__ uint32_t s_tile, s_prefix[2], s_last;
    pdl_trigger();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = A.t_ntri[i];
    for (int i = threadIdx.x; i < 256 * 16 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(s_tris)[i] = reinterpret_cast<const uint4 *>(A.t_tris)[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t bar0 = smem_u32(&s_bar[warp][0]), bar1 = smem_u32(&s_bar[warp][1]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();
    stage_stamp(A.stamp);
    __syncthreads();
    const MeshDims &D = A.D;
    const uint32_t nblk = *A.nblk;
    const uint32_t stride = gridDim.x * kBlkWarps;
    const bool fused = A.scan_state != nullptr;  // the segment scan runs inside this pass (no k_scan_seg launch)
    uint32_t phases = 0u;  // bit b = parity to wait for on buffer b
    uint32_t it = blockIdx.x * kBlkWarps + warp;
    int cur = 0;
    // the first stencil copy goes out before the scan phase: it is in flight while this CTA scans
    if (it < nblk && lane == 0) blk_issue(&tmap, reinterpret_cast<float *>(s_buf[warp][0]), bar0, D, blk_decode(D, A.blklist[it]));
    if (fused) blk_scan_phase(A, s_w, &s_tile, s_prefix);
    for (; it < nblk; it += stride, cur ^= 1) {
        const uint32_t bid = A.blklist[it];
        const BlkPos b = blk_decode(D, bid);
        if (it + stride < nblk && lane == 0)
            blk_issue(&tmap, reinterpret_cast<float *>(s_buf[warp][cur ^ 1]), cur ? bar0 : bar1, D, blk_decode(D, A.blklist[it + stride]));
        uint32_t rown[2], rowbase[2];
        blk_emit_rows(A, b, lane, fused, rown, rowbase);
        const uint32_t cm = A.childmask ? A.childmask[bid] : 0xffu;
        blk_wait(cur ? bar1 : bar0, (phases >> cur) & 1u);  // (consumed even when the block turns out empty: the barrier phase must advance)
        phases ^= 1u << cur;
        blk_emit_tris(A, b, cm, reinterpret_cast<const float *>(s_buf[warp][cur]), s_ntri, s_tris, s_list[warp], lane, rown, rowbase);
        __syncwarp();  // everybody is done with buf before it is refilled two iterations later
    }
    blk_finish(A, &s_last);
}

}  // namespace gsdfk
