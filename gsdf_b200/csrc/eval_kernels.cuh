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

template <int P, class Gen, bool EXT>
__global__ void __launch_bounds__(kEvalThreads, GSDF_EVAL_MINB) k_eval(ProgView pv, Gen gen) {
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
#ifdef GSDF_LOCKSTEP
        // every thread of the tile runs the program (barriers inside); threads past the end redo the last item
        const uint64_t wc = w < nwork ? w : nwork - 1;
        m.init(dstk, pstk, blockDim.x);
#ifdef GSDF_RXY
        m.rxy = pstk + (size_t)pv.pslots * 3 * P * blockDim.x;
#endif
        gen.load(wc, m.px, m.py, m.pz);
        run_program<P, EXT>(m, prog, aux);
        if (w < nwork) gen.store(w, m.top);
#else
        if (w >= nwork) continue;
        m.init(dstk, pstk, blockDim.x);
#ifdef GSDF_RXY
        m.rxy = pstk + (size_t)pv.pslots * 3 * P * blockDim.x;
#endif
        gen.load(w, m.px, m.py, m.pz);
        run_program<P, EXT>(m, prog, aux);
        gen.store(w, m.top);
#endif
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
