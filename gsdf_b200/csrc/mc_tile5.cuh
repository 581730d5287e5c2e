// mc_tile5.cuh -- marching cubes on 128 x 8 x 4-cell tiles (default since round 2).
//
// A tile's 5 x 9 x 132 corner stencil (23.8 KB) is fetched by ONE 3-D tensor copy (cp.async.bulk.tensor.3d -> SASS UTMALDG)
// into shared memory; four cell layers are classified from five corner planes, so every lattice plane is read 1.25 times
// per pass instead of twice (one layer per tile), and a CTA waits for a quarter as many copies. Tiles whose prune blocks
// are all empty never issue the copy. Both passes walk the same tiles:
//   k_mc_count5   classification only: a triangle count per 32-cell row segment (and the case bytes for parity checks);
//                 no work list, no atomics, no CTA-wide synchronisation besides the stencil hand-over
//   (k_scan_lookback turns the counts into offsets in FlatRenderer cell order)
//   k_mc_emit5    loads the tile again (an L2 hit), classifies again from shared memory (20 sign tests per 4 cells) and
//                 places the triangles of every non-empty segment; corner values come from the shared-memory stencil,
//                 so the vertex loop has no dependent global loads left, and a warp's vertices leave through a small
//                 shared-memory transpose as three fully coalesced 128-byte stores.
// A warp owns one cell row of the tile per layer; lane l owns the 4 cells of prune block l of that row (bit l of the
// row's prune word is its verdict). Output is identical to k_mc_count / k_mc_emit (mc_kernels.cuh).
#pragma once
#include "mc_kernels.cuh"

namespace gsdfk {

constexpr int kT5Layers = 4, kBox5Z = 5;
constexpr uint32_t kTile5Bytes = kBox5Z * kBoxY * kBoxX * 4;

struct Tile5Ctx {
    uint32_t ntx, nty, ntz, ntiles;
};
__device__ __forceinline__ Tile5Ctx tile5_ctx(const MeshDims &D) {
    Tile5Ctx c;
    c.ntx = (uint32_t)(D.nsx + 3) / 4;
    c.nty = (uint32_t)(D.ny + kTileY - 1) / kTileY;
    c.ntz = (uint32_t)(D.cz1 - D.cz0 + kT5Layers - 1) / kT5Layers;
    c.ntiles = c.ntx * c.nty * c.ntz;
    return c;
}

// Prune words of the tile: w[k][r] = bit row of block layer of cell layer k (k = 0..3) and block row of cell row cy0 + 4r
// (r = 0, 1); 0xffffffff without pruning, 0 for rows / layers outside the slab. Returns their OR (CTA-uniform).
__device__ __forceinline__ uint32_t tile5_words(const MCArgs &A, uint32_t tx, int cy0, int czb, int nl, uint32_t (&w)[kT5Layers][2]) {
    const MeshDims &D = A.D;
    uint32_t any = 0u;
#pragma unroll
    for (int k = 0; k < kT5Layers; k++) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            uint32_t v = 0u;
            const int by = (cy0 >> 2) + r;
            if (k < nl && by < D.nby) {
                v = 0xffffffffu;
                if (A.mbits) v = A.mbits[((size_t)(((czb + k) >> 2) - D.bz0) * D.nby + by) * D.nwx + tx];
            }
            w[k][r] = v;
            any |= v;
        }
    }
    return any;
}

// Sign masks and reject bits of the warp's row for the five planes of the tile, then the four cube-case indices of the
// lane's cells in layer k (one byte each; cell cx0 + c in byte c). Corner order flatrenderer.go:222-233, reject rule
// :218-220, case index marchcubes.go:39-44.
struct Row5 {
    uint32_t sm[kBox5Z][2];  // 5-bit sign masks: plane p, row y + r
    uint32_t rej[kT5Layers]; // 4-bit: |corner 0| > cubeDiag for the lane's 4 cells in layer k
};
__device__ __forceinline__ void row5_load(const float (*s)[kBoxY][kBoxX], int warp, int lane, int nl, float cubeDiag, Row5 &R) {
    const int lx = 4 * lane;
#pragma unroll
    for (int p = 0; p < kBox5Z; p++) {
        if (p > nl) { R.sm[p][0] = R.sm[p][1] = 0u; continue; }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const float4 q = *reinterpret_cast<const float4 *>(&s[p][warp + r][lx]);
            float e = __shfl_down_sync(0xffffffffu, q.x, 1);
            if (lane == 31) e = s[p][warp + r][kTileX];
            R.sm[p][r] = (q.x < 0.f ? 1u : 0u) | (q.y < 0.f ? 2u : 0u) | (q.z < 0.f ? 4u : 0u) | (q.w < 0.f ? 8u : 0u) | (e < 0.f ? 16u : 0u);
            if (r == 0 && p < kT5Layers)
                R.rej[p] = (fabsf(q.x) > cubeDiag ? 1u : 0u) | (fabsf(q.y) > cubeDiag ? 2u : 0u) | (fabsf(q.z) > cubeDiag ? 4u : 0u) | (fabsf(q.w) > cubeDiag ? 8u : 0u);
        }
    }
}
__device__ __forceinline__ uint32_t row5_cases(const Row5 &R, int k, bool act, int cx0, int nx) {
    uint32_t cases4 = 0u;
    if (!act) return 0u;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t idx = ((R.sm[k][0] >> c) & 3u) | (rev2((R.sm[k][1] >> c) & 3u) << 2) | (((R.sm[k + 1][0] >> c) & 3u) << 4) | (rev2((R.sm[k + 1][1] >> c) & 3u) << 6);
        if (cx0 + c >= nx || ((R.rej[k] >> c) & 1u)) idx = 0u;
        cases4 |= idx << (8 * c);
    }
    return cases4;
}

__device__ __forceinline__ void tile5_issue(const CUtensorMap *tmap, void *s_dst, uint32_t bar, int x, int y, int z) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kTile5Bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(s_dst)),
                 "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar)
                 : "memory");
}
// One thread polls the mbarrier, the others wait for it on the CTA barrier: a spinning try_wait loop is issued instructions
// (ncu counted more warp instructions in the polls of 256 threads than in the classification itself), bar.sync is not.
__device__ __forceinline__ void tile5_wait(uint32_t bar, uint32_t phase) {
    if (threadIdx.x == 0) asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "T5_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra T5_DONE;\n\t"
        "bra T5_LOOP;\n\t"
        "T5_DONE:\n\t"
        "}" ::"r"(bar), "r"(phase)
        : "memory");
    __syncthreads();
}

__global__ void __launch_bounds__(256, 8) k_mc_count5(const __grid_constant__ CUtensorMap tmap, MCArgs A) {
    __shared__ __align__(128) float s_tile[kBox5Z][kBoxY][kBoxX];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint8_t s_ntri[256];
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
    const Tile5Ctx T = tile5_ctx(D);
    uint32_t phase = 0;
    for (uint32_t tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const uint32_t tx = tile % T.ntx, ty = (tile / T.ntx) % T.nty, tz = tile / (T.ntx * T.nty);
        const int cy0 = (int)ty * kTileY, czb = D.cz0 + (int)tz * kT5Layers;
        const int nl = min(kT5Layers, D.cz1 - czb);
        uint32_t w[kT5Layers][2];
        const uint32_t any = tile5_words(A, tx, cy0, czb, nl, w);
        const int cy = cy0 + warp;
        const int nsg = cy < D.ny ? min(4, D.nsx - (int)tx * 4) : 0;
        const int cx0 = (int)tx * kTileX + 4 * lane;
        if (any == 0u) {  // CTA-uniform: no kept block in the tile
#pragma unroll
            for (int k = 0; k < kT5Layers; k++) {
                if (k >= nl) break;
                const uint32_t r = (uint32_t)(czb + k - D.cz0) * (uint32_t)D.ny + (uint32_t)cy;
                if (lane < nsg) A.segcount[r * (uint32_t)D.nsx + tx * 4u + lane] = 0u;
                if (A.cases && cy < D.ny) {
#pragma unroll
                    for (int c = 0; c < 4; c++) if (cx0 + c < D.nx) A.cases[(size_t)r * D.nx + cx0 + c] = 0;
                }
            }
            continue;
        }
        if (threadIdx.x == 0) tile5_issue(&tmap, &s_tile[0][0][0], bar, (int)tx * kTileX, cy0, (int)tz * kT5Layers);
        tile5_wait(bar, phase);
        phase ^= 1u;
        Row5 R;
        row5_load(s_tile, warp, lane, nl, A.cubeDiag, R);
#pragma unroll
        for (int k = 0; k < kT5Layers; k++) {
            if (k >= nl) break;
            const uint32_t word = cy < D.ny ? (((cy >> 2) == (cy0 >> 2)) ? w[k][0] : w[k][1]) : 0u;
            const uint32_t cases4 = row5_cases(R, k, ((word >> lane) & 1u) != 0u, cx0, D.nx);
            uint32_t mine = 0u;
            // some byte outside {0, 255}  <=>  (x ^ 0xff in every byte whose low bit is set) != 0
            const uint32_t mixed = cases4 ^ ((cases4 & 0x01010101u) * 255u);
            if (__ballot_sync(0xffffffffu, mixed != 0u) != 0u) {
                uint32_t n = (uint32_t)s_ntri[cases4 & 0xffu] + s_ntri[(cases4 >> 8) & 0xffu] + s_ntri[(cases4 >> 16) & 0xffu] + s_ntri[cases4 >> 24];
                n += __shfl_xor_sync(0xffffffffu, n, 1);
                n += __shfl_xor_sync(0xffffffffu, n, 2);
                n += __shfl_xor_sync(0xffffffffu, n, 4);   // every lane of an 8-lane group holds its segment's total
                mine = __shfl_sync(0xffffffffu, n, (lane & 3) * 8);  // lane sgm (< 4) keeps the count of segment sgm
            }
            const uint32_t r = (uint32_t)(czb + k - D.cz0) * (uint32_t)D.ny + (uint32_t)cy;
            if (lane < nsg) A.segcount[r * (uint32_t)D.nsx + tx * 4u + lane] = mine;
            if (A.cases && cy < D.ny) {
#pragma unroll
                for (int c = 0; c < 4; c++) if (cx0 + c < D.nx) A.cases[(size_t)r * D.nx + cx0 + c] = (uint8_t)(cases4 >> (8 * c));
            }
        }
        __syncthreads();  // every warp is done reading s_tile before the next tile's copy may land
    }
}

// Pass 2 on the same tiles. segcount[] holds exclusive triangle offsets now.
__global__ void __launch_bounds__(256, 6) k_mc_emit5(const __grid_constant__ CUtensorMap tmap, MCArgs A) {
    __shared__ __align__(128) float s_tile[kBox5Z][kBoxY][kBoxX];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint8_t s_ntri[256];
    __shared__ __align__(16) int8_t s_tris[256 * 16];
    __shared__ float s_out[8][96];  // per warp: 32 vertices x 3 floats, transposed into three coalesced stores
    pdl_trigger();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = A.t_ntri[i];
    for (int i = threadIdx.x; i < 256 * 16 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(s_tris)[i] = reinterpret_cast<const uint4 *>(A.t_tris)[i];
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();
    stage_stamp(A.stamp);
    __syncthreads();
    const MeshDims &D = A.D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Tile5Ctx T = tile5_ctx(D);
    const float rr = A.res;
    float *so = s_out[warp];
    uint32_t phase = 0;
    for (uint32_t tile = blockIdx.x; tile < T.ntiles; tile += gridDim.x) {
        const uint32_t tx = tile % T.ntx, ty = (tile / T.ntx) % T.nty, tz = tile / (T.ntx * T.nty);
        const int cy0 = (int)ty * kTileY, czb = D.cz0 + (int)tz * kT5Layers;
        const int nl = min(kT5Layers, D.cz1 - czb);
        uint32_t w[kT5Layers][2];
        if (tile5_words(A, tx, cy0, czb, nl, w) == 0u) continue;  // CTA-uniform
        if (threadIdx.x == 0) tile5_issue(&tmap, &s_tile[0][0][0], bar, (int)tx * kTileX, cy0, (int)tz * kT5Layers);
        tile5_wait(bar, phase);
        phase ^= 1u;
        const int cy = cy0 + warp;
        const int cx0 = (int)tx * kTileX + 4 * lane;
        Row5 R;
        row5_load(s_tile, warp, lane, nl, A.cubeDiag, R);
        const float py0 = A.oy + (float)cy * rr, py1 = py0 + rr;
#pragma unroll
        for (int k = 0; k < kT5Layers; k++) {
            if (k >= nl) break;
            const uint32_t word = cy < D.ny ? (((cy >> 2) == (cy0 >> 2)) ? w[k][0] : w[k][1]) : 0u;
            const uint32_t cases4 = row5_cases(R, k, ((word >> lane) & 1u) != 0u, cx0, D.nx);
            const uint32_t mixed = cases4 ^ ((cases4 & 0x01010101u) * 255u);
            const uint32_t live = __ballot_sync(0xffffffffu, mixed != 0u);  // lanes whose cells may hold triangles
            if (live == 0u) continue;                                       // warp-uniform
            const int cz = czb + k;
            const uint32_t r = (uint32_t)(cz - D.cz0) * (uint32_t)D.ny + (uint32_t)cy;
            const uint32_t s0 = r * (uint32_t)D.nsx + tx * 4u;
            const float pz0 = A.oz + (float)cz * rr, pz1 = pz0 + rr;
#pragma unroll 1
            for (int sgm = 0; sgm < 4; sgm++) {
                if (((live >> (8 * sgm)) & 0xffu) == 0u) continue;  // warp-uniform: nothing in this 32-cell segment
                // lane = cell of the segment: its case byte sits in lane 8 sgm + lane/4, byte lane%4
                const int index = (int)((__shfl_sync(0xffffffffu, cases4, 8 * sgm + (lane >> 2)) >> (8 * (lane & 3))) & 0xffu);
                const uint32_t n = s_ntri[index];
                const uint32_t incl = warp_incl_scan(n);
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                if (total == 0u) continue;
                const uint64_t obase = (uint64_t)A.segcount[s0 + sgm];
                const int x0l = 32 * sgm;                       // first cell of the segment inside the tile
                const int x0 = (int)tx * kTileX + x0l;
                for (uint32_t ibase = 0; ibase < 3u * total; ibase += 32) {  // warp-uniform trip count: shuffles use all lanes
                    const uint32_t item = ibase + lane;
                    const bool alive = item < 3u * total;
                    const uint32_t tri = alive ? item / 3u : total - 1u, j = alive ? item - 3u * tri : 0u;
                    int lo = 0;  // owner = first lane whose inclusive count exceeds tri (binary search by shuffle)
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const uint32_t probe = __shfl_sync(0xffffffffu, incl, lo + step - 1);
                        if (probe <= tri) lo += step;
                    }
                    const int owner = lo;
                    const uint32_t oincl = __shfl_sync(0xffffffffu, incl, owner);
                    const uint32_t on = __shfl_sync(0xffffffffu, n, owner);
                    const int oindex = __shfl_sync(0xffffffffu, index, owner);
                    float3 q = make_float3(0.f, 0.f, 0.f);
                    if (alive) {
                        const uint32_t kk = tri - (oincl - on);
                        const float px0 = A.ox + (float)(x0 + owner) * rr, px1 = px0 + rr;
                        // marchcubes.go:64-68: vertex j of the triangle is points[table[3k + 2 - j]]
                        const int e = s_tris[16 * oindex + 3 * (int)kk + 2 - (int)j];
                        // edge -> corner pair (marchcubes.go:101-114), packed 4 bits per edge
                        const int ca = (int)((0x321076543210ull >> (4 * e)) & 0xf), cb = (int)((0x765447650321ull >> (4 * e)) & 0xf);
                        const float3 pa = make_float3((((ca + 1) >> 1) & 1) ? px1 : px0, ((ca >> 1) & 1) ? py1 : py0, (ca >> 2) ? pz1 : pz0);
                        const float3 pb = make_float3((((cb + 1) >> 1) & 1) ? px1 : px0, ((cb >> 1) & 1) ? py1 : py0, (cb >> 2) ? pz1 : pz0);
                        // corner c of cell (x0 + owner): x bit ((c+1)>>1)&1, y bit (c>>1)&1, z bit c>>2 (flatrenderer.go:222-233)
                        const float va = s_tile[k + (ca >> 2)][warp + ((ca >> 1) & 1)][x0l + owner + (((ca + 1) >> 1) & 1)];
                        const float vb = s_tile[k + (cb >> 2)][warp + ((cb >> 1) & 1)][x0l + owner + (((cb + 1) >> 1) & 1)];
                        q = mc_interp(pa, pb, va, vb);
                    }
                    // 32 vertices x 3 floats = 96 consecutive floats of the output: out through a transpose, three coalesced stores
                    const uint64_t o0 = obase * 9u + 3u * ibase;                 // first float of this batch
                    const uint32_t nfl = min(96u, 9u * total - 3u * ibase);     // floats of this batch that exist
                    __syncwarp();
                    so[3 * lane] = q.x; so[3 * lane + 1] = q.y; so[3 * lane + 2] = q.z;
                    __syncwarp();
                    if (o0 + nfl > A.tri_capacity * 9u) {
                        if (lane == 0) *A.overflow = 1u;
                    } else {
#pragma unroll
                        for (int t = 0; t < 3; t++) {
                            const uint32_t f = (uint32_t)t * 32u + (uint32_t)lane;
                            if (f < nfl) A.tris[o0 + f] = so[f];
                        }
                    }
                }
            }
        }
        __syncthreads();  // every warp is done reading s_tile before the next tile's copy may land
    }
    if (A.fin_ctr) {  // the last CTA to get here ends the render (k_finish_render's work, without its launch)
        __shared__ uint32_t s_last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = atomicAdd(A.fin_done, 1u) == gridDim.x - 1 ? 1u : 0u;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            finish_render_cta(A.fin_ctr, A.fin_hctr, A.fin_nctr, A.fin_scanstate, A.fin_nstate, A.fin_dstamp, A.fin_hstamp, A.fin_nstamp);
        }
    }
}

}  // namespace gsdfk
