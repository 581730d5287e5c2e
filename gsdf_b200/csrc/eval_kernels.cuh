// eval_kernels.cuh -- the interpreter kernels (compiled once, in eval.cu).
//
//   k_eval<P,Gen>      persistent CTAs pull kEvalThreads-item tiles from an atomic counter; each thread interprets the node
//                      program at P points produced by a generator functor (AoS point lists, the dense lattice, a
//                      compacted quad list, prune-cube centres, image rows) and hands the distances to its sink.
//   k_eval_stream      gleval.SDF3.Evaluate on device-resident point lists, bulk-async double-buffered position tiles.
//
// The node program (+ side buffer when it fits) is staged into shared memory once per CTA by a 1-D bulk async copy
// (cp.async.bulk.shared::cluster.global, completion on an mbarrier: SASS UBLKCP / SYNCS).
#pragma once
#include "generators.cuh"
#include "interp.cuh"

namespace gsdfk {

// How a tile's threads run the node program: the interpreter loop (RunInterp), or the straight-line specialisation of one
// program that jit.cu generates and compiles at run time (RunSpecial there).
template <int P, bool EXT>
struct RunInterp {
    __device__ __forceinline__ void operator()(Machine<P> &m, const uint4 *__restrict__ prog, const float4 *__restrict__ aux) const { run_program<P, EXT>(m, prog, aux); }
};

// The body of k_eval, shared with the run-time compiled kernels. ALL_RUN: every thread of a tile runs the program (barriers or
// CTA-wide guard votes inside); threads past the end redo the last item.
template <int P, bool ALL_RUN, class Gen, class Run>
__device__ __forceinline__ void eval_body(ProgView pv, Gen gen, Run run) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t stage = smem_stage_bytes(pv);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + stage);
    volatile uint32_t *s_tile = reinterpret_cast<volatile uint32_t *>(smem + stage + 8);
    pdl_trigger();
    bulk_stage(smem, pv.g_prog, stage, bar);  // the program was uploaded before the chain started: safe ahead of pdl_wait
    const uint4 *prog = reinterpret_cast<const uint4 *>(smem);
    const float4 *aux = pv.stage_aux ? reinterpret_cast<const float4 *>(smem + pv.prog_bytes)
                                     : reinterpret_cast<const float4 *>(reinterpret_cast<const uint8_t *>(pv.g_prog) + pv.prog_bytes);
    float *dstk = reinterpret_cast<float *>(smem + stage + 16u) + threadIdx.x;
    float *pstk = dstk + (size_t)pv.dslots * P * blockDim.x;

    Machine<P> m;
    pdl_wait();
    stage_stamp(pv.stamp);
    const uint64_t nwork = gen.work_items();
    // the first tile of a CTA is its own index (no round trip to the counter: thin slabs are one tile per CTA and latency
    // bound); further tiles come from the launch's scheduler counter, offset by the grid size
    for (bool first = true;; first = false) {
        uint64_t w;
        if (first) {
            w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        } else {
            if ((uint64_t)gridDim.x * blockDim.x >= nwork) break;  // the first tiles covered the work: nothing to fetch
            if (threadIdx.x == 0) *s_tile = gridDim.x + atomicAdd(pv.sched, 1u);
            __syncthreads();
            w = (uint64_t)(*s_tile) * blockDim.x + threadIdx.x;
            __syncthreads();
        }
        if (w - threadIdx.x >= nwork) break;
        if constexpr (Gen::kTileSkip) {  // generator-level CTA-uniform skip of a whole tile (GenDC: cubes outside this rank's region)
            if (__syncthreads_and(gen.dead(w < nwork ? w : nwork - 1))) {
                if (w < nwork) gen.store_dead(w);
                continue;
            }
        }
        if constexpr (ALL_RUN) {
            const uint64_t wc = w < nwork ? w : nwork - 1;
            m.init(dstk, pstk, blockDim.x);
#ifdef GSDF_RXY
            m.rxy = pstk + (size_t)pv.pslots * 3 * P * blockDim.x;
#endif
            gen.load(wc, m.px, m.py, m.pz);
            run(m, prog, aux);
            if (w < nwork) gen.store(w, m.top);
        } else {
            if (w >= nwork) continue;
            m.init(dstk, pstk, blockDim.x);
#ifdef GSDF_RXY
            m.rxy = pstk + (size_t)pv.pslots * 3 * P * blockDim.x;
#endif
            gen.load(w, m.px, m.py, m.pz);
            run(m, prog, aux);
            gen.store(w, m.top);
        }
    }
    // the last CTA to leave re-arms the scheduler for the next launch
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(pv.sched + 1, 1u) == gridDim.x - 1) {
            pv.sched[0] = 0u;
            pv.sched[1] = 0u;
        }
    }
}

// The same for kernels whose warps are independent (the run-time compiled ones: no lockstep barriers, guards vote per warp):
// a tile is the 32 work items of ONE warp, fetched by the warp itself, so the tail of the launch is balanced at warp
// granularity and nothing inside the tile loop waits for another warp.
template <int P, class Gen, class Run>
__device__ __forceinline__ void eval_body_warp(ProgView pv, Gen gen, Run run) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t stage = smem_stage_bytes(pv);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + stage);
    pdl_trigger();
    bulk_stage(smem, pv.g_prog, stage, bar);
    const uint4 *prog = reinterpret_cast<const uint4 *>(smem);
    const float4 *aux = pv.stage_aux ? reinterpret_cast<const float4 *>(smem + pv.prog_bytes)
                                     : reinterpret_cast<const float4 *>(reinterpret_cast<const uint8_t *>(pv.g_prog) + pv.prog_bytes);
    float *dstk = reinterpret_cast<float *>(smem + stage + 16u) + threadIdx.x;
    float *pstk = dstk + (size_t)pv.dslots * P * blockDim.x;
    const uint32_t lane = threadIdx.x & 31u, wpb = blockDim.x >> 5, nwarps = gridDim.x * wpb;

    Machine<P> m;
    pdl_wait();
    stage_stamp(pv.stamp);
    const uint64_t nwork = gen.work_items();
    for (bool first = true;; first = false) {
        uint32_t tile = blockIdx.x * wpb + (threadIdx.x >> 5);  // a warp's first tile is its own index: no counter round trip
        if (!first) {
            // (when the first tiles cover the work -- the centre pass, thin slabs -- there is nothing to fetch: thousands of
            // same-address atomics, or reads, at the end of a one-tile-per-warp launch cost microseconds)
            if ((uint64_t)nwarps * 32u >= nwork) break;
            if (lane == 0) tile = nwarps + atomicAdd(pv.sched, 1u);
            tile = __shfl_sync(0xffffffffu, tile, 0);
        }
        const uint64_t w = (uint64_t)tile * 32u + lane;
        if (w - lane >= nwork) break;
        const uint64_t wc = w < nwork ? w : nwork - 1;  // lanes past the end redo the last item (warp-wide votes inside)
        if constexpr (Gen::kTileSkip) {
            if (__all_sync(0xffffffffu, gen.dead(wc))) {
                if (w < nwork) gen.store_dead(w);
                continue;
            }
        }
        m.init(dstk, pstk, blockDim.x);
#ifdef GSDF_RXY
        m.rxy = pstk + (size_t)pv.pslots * 3 * P * blockDim.x;
#endif
        gen.load(wc, m.px, m.py, m.pz);
        run(m, prog, aux);
        if (w < nwork) gen.store(w, m.top);
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // the last CTA to leave re-arms the scheduler for the next launch
        __threadfence();
        if (atomicAdd(pv.sched + 1, 1u) == gridDim.x - 1) {
            pv.sched[0] = 0u;
            pv.sched[1] = 0u;
        }
    }
}

#ifdef GSDF_LOCKSTEP
constexpr bool kInterpAllRun = true;
#else
constexpr bool kInterpAllRun = false;
#endif
template <int P, class Gen, bool EXT>
__global__ void __launch_bounds__(kEvalThreads, GSDF_EVAL_MINB) k_eval(ProgView pv, Gen gen) {
    eval_body<P, kInterpAllRun>(pv, gen, RunInterp<P, EXT>());
}

// ---------------------------------------------------------------------------------------------- prune levels 3 + 2
// See PruneFine (generators.cuh). Work items are level-3 cube rows padded to whole warps, exactly as GenCenters'.
// Shared memory: [prog (+aux)] [mbarrier + tile slot, 16 bytes] [dstack] [pstack] [radius] (one point per thread), then
// [parent slots: blockDim u32] [child masks: blockDim u32] [warp sums: 32 u32] [kept count]
__host__ __device__ inline uint32_t prune_fine_extra_bytes(int threads) { return (uint32_t)threads * 8u + 32u * 4u + 16u; }

template <bool EXT>
__global__ void __launch_bounds__(kEvalThreads, GSDF_EVAL_MINB) k_prune_fine(ProgView pv, PruneFine g) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t stage = smem_stage_bytes(pv);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + stage);
    volatile uint32_t *s_tile = reinterpret_cast<volatile uint32_t *>(smem + stage + 8);
    pdl_trigger();
    bulk_stage(smem, pv.g_prog, stage, bar);
    const uint4 *prog = reinterpret_cast<const uint4 *>(smem);
    const float4 *aux = pv.stage_aux ? reinterpret_cast<const float4 *>(smem + pv.prog_bytes)
                                     : reinterpret_cast<const float4 *>(reinterpret_cast<const uint8_t *>(pv.g_prog) + pv.prog_bytes);
    float *dstk = reinterpret_cast<float *>(smem + stage + 16u) + threadIdx.x;
    float *pstk = dstk + (size_t)pv.dslots * blockDim.x;
    uint32_t *s_par = reinterpret_cast<uint32_t *>(smem + smem_total_bytes<1>(pv, blockDim.x));
    uint32_t *s_cm = s_par + blockDim.x;
    uint32_t *s_wsum = s_cm + blockDim.x;
    volatile uint32_t *s_kept = s_wsum + 32;
    const PruneLevel &L = g.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t rowlen = (uint32_t)L.nwx * 32u;
    const uint64_t nwork = (uint64_t)rowlen * L.ncy * L.ncz;

    Machine<1> m;
    pdl_wait();
    stage_stamp(pv.stamp);
    for (bool first = true;; first = false) {
        uint64_t w0;  // first work item of the tile
        if (first) {
            w0 = (uint64_t)blockIdx.x * blockDim.x;
        } else {
            if ((uint64_t)gridDim.x * blockDim.x >= nwork) break;  // the first tiles covered the work: nothing to fetch
            if (threadIdx.x == 0) *s_tile = gridDim.x + atomicAdd(pv.sched, 1u);
            __syncthreads();
            w0 = (uint64_t)(*s_tile) * blockDim.x;
            __syncthreads();
        }
        if (w0 >= nwork) break;
        const uint64_t w = w0 + threadIdx.x;
        const bool inrange = w < nwork;  // whole warps: nwork and the tile size are multiples of 32
        const uint64_t wc = inrange ? w : nwork - 1;
        int cx, cy, cz;
        {
            const uint32_t row = (uint32_t)(wc / rowlen);
            cx = (int)(wc - (uint64_t)row * rowlen);
            cy = (int)(row % (uint32_t)L.ncy);
            cz = (int)(row / (uint32_t)L.ncy);
        }
        bool alive = inrange && cx < L.ncx;
        if (alive && g.P.bits) {
            const int px = cx >> g.shift, py = cy >> g.shift, pz = ((L.cz0 + cz) >> g.shift) - g.P.cz0;
            alive = (g.P.bits[((size_t)pz * g.P.ncy + py) * g.P.nwx + (px >> 5)] >> (px & 31)) & 1u;
        }
        uint32_t *row2 = g.bits2 + (((size_t)2 * cz) * (2 * L.ncy) + 2 * cy) * (2 * L.nwx) + 2 * (cx >> 5);  // (dz, dy) = (0, 0)
        const size_t dy2 = (size_t)2 * L.nwx, dz2 = (size_t)2 * L.ncy * dy2;
        if (__syncthreads_and(!alive)) {  // every cube of the tile has a pruned parent (or is padding): nothing to evaluate
            if (inrange && lane == 0) {
                L.bits[w >> 5] = 0u;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t *r = row2 + (q >> 1) * dz2 + (q & 1) * dy2;
                    r[0] = 0u; r[1] = 0u;
                }
            }
            continue;
        }
        // ---- level 3: the cube's own centre (padding lanes and lanes behind a pruned parent repeat a valid cube)
        m.init(dstk, pstk, blockDim.x);
#ifdef GSDF_RXY
        m.rxy = pstk + (size_t)pv.pslots * 3 * blockDim.x;
#endif
        {
            const int ccx = min(cx, L.ncx - 1);
            m.px[0] = (g.ox + (float)(L.w * ccx) * g.res) + L.half;
            m.py[0] = (g.oy + (float)(L.w * cy) * g.res) + L.half;
            m.pz[0] = (g.oz + (float)(L.w * (L.cz0 + cz)) * g.res) + L.half;
        }
        run_program<1, EXT>(m, prog, aux);
        const bool keep = alive && !(fabsf(m.top[0]) >= L.maxDist);
        // ---- compact the survivors of the tile
        const uint32_t kb = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_wsum[warp] = (uint32_t)__popc(kb);
        s_cm[threadIdx.x] = 0u;
        __syncthreads();
        uint32_t slot = (uint32_t)__popc(kb & ((1u << lane) - 1u));
        for (int ww = 0; ww < warp; ww++) slot += s_wsum[ww];
        if (keep) s_par[slot] = threadIdx.x;
        if (threadIdx.x == 0) {
            uint32_t k = 0;
            for (int ww = 0; ww < nwarps; ww++) k += s_wsum[ww];
            *s_kept = k;
        }
        __syncthreads();
        const uint32_t nchild = 8u * *s_kept;
        // ---- level 2: the eight children of every survivor, blockDim.x at a time
        uint32_t nev2 = 0;
        for (uint32_t base = 0; base < nchild; base += blockDim.x) {  // CTA-uniform trip count
            const uint32_t item = base + threadIdx.x;
            const bool valid = item < nchild;
            const uint32_t it = valid ? item : nchild - 1u;
            const uint32_t par = s_par[it >> 3], c = it & 7u;
            const uint64_t wp = w0 + par;
            const uint32_t prow = (uint32_t)(wp / rowlen);
            const int pcx = (int)(wp - (uint64_t)prow * rowlen), pcy = (int)(prow % (uint32_t)L.ncy), pcz = (int)(prow / (uint32_t)L.ncy);
            const int c2x = 2 * pcx + (int)(c & 1u), c2y = 2 * pcy + (int)((c >> 1) & 1u), c2z = 2 * (L.cz0 + pcz) + (int)(c >> 2);
            const bool exists = 2 * c2x < g.nx && 2 * c2y < g.ny && 2 * c2z < g.nz;
            m.init(dstk, pstk, blockDim.x);
#ifdef GSDF_RXY
            m.rxy = pstk + (size_t)pv.pslots * 3 * blockDim.x;
#endif
            m.px[0] = (g.ox + (float)(2 * c2x) * g.res) + g.half2;
            m.py[0] = (g.oy + (float)(2 * c2y) * g.res) + g.half2;
            m.pz[0] = (g.oz + (float)(2 * c2z) * g.res) + g.half2;
            run_program<1, EXT>(m, prog, aux);
            if (valid && exists) {
                nev2++;
                if (!(fabsf(m.top[0]) >= g.maxDist2)) atomicOr(&s_cm[par], 1u << c);
            }
        }
        __syncthreads();
        // ---- publish: level-3 word, child masks, level-2 rows (E, O words), counters
        const uint32_t cm = s_cm[threadIdx.x];  // zero unless this thread's cube survived with children
        const uint32_t word3 = __ballot_sync(0xffffffffu, cm != 0u);
        const uint32_t nlive = (uint32_t)__popc(__ballot_sync(0xffffffffu, alive));
        uint32_t n2 = (uint32_t)__popc(cm);
        uint32_t ne = nev2;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { n2 += __shfl_xor_sync(0xffffffffu, n2, o); ne += __shfl_xor_sync(0xffffffffu, ne, o); }
        if (inrange) {
            if (cx < L.ncx) g.childmask[((size_t)cz * L.ncy + cy) * L.ncx + cx] = (uint8_t)cm;
#pragma unroll
            for (int q = 0; q < 4; q++) {  // q = dy + 2 dz
                const uint32_t e = __ballot_sync(0xffffffffu, (cm >> (2 * q)) & 1u), o = __ballot_sync(0xffffffffu, (cm >> (2 * q + 1)) & 1u);
                if (lane == 0) {
                    uint32_t *r = row2 + (q >> 1) * dz2 + (q & 1) * dy2;
                    r[0] = e; r[1] = o;
                }
            }
            if (lane == 0) {
                L.bits[w >> 5] = word3;
                if (word3 && g.kept) atomicAdd(g.kept, (uint32_t)__popc(word3));
                if (n2 && g.kept2) atomicAdd(g.kept2, n2);
                if (nlive + ne) atomicAdd(g.evals, nlive + ne);
            }
        } else if (ne && lane == 0) {
            atomicAdd(g.evals, ne);
        }
        __syncthreads();  // the shared lists are reused by the next tile
    }
    if (threadIdx.x == 0) {  // the last CTA to leave re-arms the scheduler for the next launch
        __threadfence();
        if (atomicAdd(pv.sched + 1, 1u) == gridDim.x - 1) {
            pv.sched[0] = 0u;
            pv.sched[1] = 0u;
        }
    }
}

// gleval.SDF3.Evaluate / SDF2.Evaluate on device-resident point lists, streaming form. A tile is the AoS position block of
// blockDim.x * 4 points -- one contiguous run of global memory (24 KB for float3, 16 KB for float2) -- fetched by ONE
// 1-D bulk async copy (cp.async.bulk.shared::cluster.global -> SASS UBLKCP) into a double-buffered shared-memory stage:
// the copy of tile i+1 is in flight while tile i is interpreted, so cheap trees (sphere, box: 16 B/eval) stay on the HBM
// stream instead of alternating load and compute phases. Shared-memory reads are 3 (2) float4 per thread at a 48 (32)
// byte stride: conflict-free per quarter-warp. Tiles are dealt round-robin to the persistent CTAs (cost per tile is
// uniform); the last, partial tile takes the plain-load path. Requires 16-byte aligned pos/dist (else k_eval<GenPoints*>).
// Shared memory: [prog (+aux)] [2 mbarriers] [dstack] [pstack] [stage 0] [stage 1]
template <int DIM>
__host__ __device__ inline uint32_t stream_stage_bytes(int threads) { return (uint32_t)threads * 4u * DIM * 4u; }

template <int DIM, bool EXT>
__global__ void __launch_bounds__(kEvalThreads) k_eval_stream(ProgView pv, const float *__restrict__ pos, float *__restrict__ dist, uint64_t n) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int P = 4;
    const uint32_t stage = smem_stage_bytes(pv);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + stage);  // bar[0]: program staging, then stage 0; bar[1]: stage 1
    bulk_stage(smem, pv.g_prog, stage, bar);                     // completes phase 0 of bar[0]
    const uint4 *prog = reinterpret_cast<const uint4 *>(smem);
    const float4 *aux = pv.stage_aux ? reinterpret_cast<const float4 *>(smem + pv.prog_bytes)
                                     : reinterpret_cast<const float4 *>(reinterpret_cast<const uint8_t *>(pv.g_prog) + pv.prog_bytes);
    float *dstk = reinterpret_cast<float *>(smem + stage + 16u) + threadIdx.x;
    float *pstk = dstk + (size_t)pv.dslots * P * blockDim.x;
    const uint32_t stack_bytes = (uint32_t)blockDim.x * P * 4u * (pv.dslots + 3u * pv.pslots + kRxySlots);
    const uint32_t tile_bytes = stream_stage_bytes<DIM>(blockDim.x);
    uint8_t *buf0 = smem + ((stage + 16u + stack_bytes + 127u) & ~127u);
    const uint32_t bar_u[2] = {smem_u32(bar), smem_u32(bar + 1)};
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u[1]));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t pts_per_tile = (uint64_t)blockDim.x * P;
    const uint64_t nfull = n / pts_per_tile;               // tiles fetched by bulk copy
    const uint64_t ntiles = (n + pts_per_tile - 1) / pts_per_tile;
    uint32_t phase[2] = {1u, 0u};                          // bar[0] already went through phase 0 for the program
    auto issue = [&](uint64_t tile, int b) {               // one elected thread
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_u[b]), "r"(tile_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf0 + (size_t)b * tile_bytes)),
                     "l"(reinterpret_cast<const uint8_t *>(pos) + tile * tile_bytes), "r"(tile_bytes), "r"(bar_u[b])
                     : "memory");
    };
    uint64_t tile = blockIdx.x;
    if (threadIdx.x == 0 && tile < nfull) issue(tile, 0);
    Machine<P> m;
    int b = 0;
    for (; tile < ntiles; tile += gridDim.x, b ^= 1) {
        const uint64_t next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < nfull) issue(next, b ^ 1);   // stage b^1 was released by the barrier that ended the previous iteration
        m.init(dstk, pstk, blockDim.x);
#ifdef GSDF_RXY
        m.rxy = pstk + (size_t)pv.pslots * 3 * P * blockDim.x;
#endif
        const uint64_t i0 = tile * pts_per_tile + (uint64_t)threadIdx.x * P;
        if (tile < nfull) {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "SW_LOOP:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra SW_DONE;\n\t"
                "bra SW_LOOP;\n\t"
                "SW_DONE:\n\t"
                "}" ::"r"(bar_u[b]), "r"(phase[b])
                : "memory");
            phase[b] ^= 1u;
            const float4 *s4 = reinterpret_cast<const float4 *>(buf0 + (size_t)b * tile_bytes) + threadIdx.x * DIM;
            if (DIM == 3) {
                const float4 a = s4[0], bq = s4[1], c = s4[2];
                m.px[0] = a.x; m.py[0] = a.y; m.pz[0] = a.z; m.px[1] = a.w; m.py[1] = bq.x; m.pz[1] = bq.y;
                m.px[2] = bq.z; m.py[2] = bq.w; m.pz[2] = c.x; m.px[3] = c.y; m.py[3] = c.z; m.pz[3] = c.w;
            } else {
                const float4 a = s4[0], bq = s4[1];
                m.px[0] = a.x; m.py[0] = a.y; m.px[1] = a.z; m.py[1] = a.w; m.px[2] = bq.x; m.py[2] = bq.y; m.px[3] = bq.z; m.py[3] = bq.w;
#pragma unroll
                for (int j = 0; j < P; j++) m.pz[j] = 0.f;
            }
        } else {  // the partial tile: plain loads, points past the end repeat the last one
#pragma unroll
            for (int j = 0; j < P; j++) {
                const uint64_t i = i0 + j < n ? i0 + j : n - 1;
                m.px[j] = __ldg(pos + DIM * i); m.py[j] = __ldg(pos + DIM * i + 1); m.pz[j] = DIM == 3 ? __ldg(pos + DIM * i + 2) : 0.f;
            }
        }
        run_program<P, EXT>(m, prog, aux);
        if (i0 + P <= n) {
            reinterpret_cast<float4 *>(dist)[i0 / P] = make_float4(m.top[0], m.top[1], m.top[2], m.top[3]);
        } else {
#pragma unroll
            for (int j = 0; j < P; j++) if (i0 + j < n) dist[i0 + j] = m.top[j];
        }
        __syncthreads();  // every thread is done with stage b before it is refilled two iterations later
    }
}

}  // namespace gsdfk
