// eval.cu -- the interpreter kernels and their launchers. This is the slow translation unit (one k_eval instantiation per
// generator and per {default, EXT} interpreter); nothing else includes interp.cuh.
#include <map>
#include <tuple>

#include "eval_kernels.cuh"
#include "internal.cuh"

using namespace gsdfk;

namespace gsdfi {

namespace {

constexpr int kMaxDev = 64;

// Per-kernel, per-device launch facts. cudaFuncSetAttribute and occupancy are per device: the opt-in is made once per
// (kernel, device) to the device's maximum (so no later, smaller program can lower it under a concurrent launch) and the
// occupancy of each dynamic-shared-memory size seen is cached per device.
struct KernelDevCache {
    std::mutex mu;
    bool optin[kMaxDev] = {};
    struct Occ { uint32_t smem; int threads; int occ; };
    std::vector<Occ> occ[kMaxDev];
};

template <class Kern>
int kernel_occupancy(KernelDevCache &c, Kern kern, int dev, uint32_t smem, int threads, int *occ_out) {
    if (dev < 0 || dev >= kMaxDev) return fail(GSDF_EINVAL, "device %d out of range", dev);
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.optin[dev]) {
        DevInfo di;
        const int rc = device_info(dev, &di);
        if (rc) return rc;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin));
        c.optin[dev] = true;
    }
    for (const auto &o : c.occ[dev])
        if (o.smem == smem && o.threads == threads) { *occ_out = o.occ; return 0; }
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    c.occ[dev].push_back({smem, threads, occ});
    *occ_out = occ;
    return 0;
}

// persistent launch: at most one resident wave of CTAs; they pull tiles from the launch's scheduler slot
// CTA size of a launch. The kernels are compiled for kEvalThreads (launch bounds, register cap) and may be launched with
// fewer threads: every lockstep barrier then spans fewer warps and an SM holds more independent CTAs. Measured on B200
// (scripts/gpu_r2_ab_cta.sh: 128 against 384 threads, flange@400 / bolt@400 / knurled@500 lattice evaluation): 67 / 63 / 508 us
// against 70 / 72 / 442 us -- short programs gain, the 40-instruction knurled tree with its large opcode bodies (atan2, two
// sincos per circular-array entry) loses to instruction-cache misses of nine free-running CTAs per SM. Hence: small CTAs
// for programs of at most kSmallProgram instructions and for the one-point-per-thread centre pass (latency bound); the
// compiled size otherwise. GSDF_EVAL_CTA=<threads> overrides (A/B).
constexpr int kSmallCta = 128;
constexpr uint32_t kSmallProgram = 24;
int eval_cta_threads(const gsdf_program *p, bool latency_bound) {
    static const int forced = getenv("GSDF_EVAL_CTA") ? atoi(getenv("GSDF_EVAL_CTA")) : 0;
    if (forced >= 32 && forced <= kEvalThreads && forced % 32 == 0) return forced;
    return (latency_bound || p->ninstr <= kSmallProgram) ? kSmallCta : kEvalThreads;
}

// CTA size of the run-time compiled kernels (no lockstep barriers inside). Short programs: 128 threads (flange@400: 47.2 us
// at 128 threads, 47.7 at 192, 49.5 at 256, 52.3 at 384). Long programs, whose straight-line code (knurled: ~130 KB) streams
// through the instruction cache whatever the CTA size: 256 (knurled@500: 449 us at 128, 433 at 256). GSDF_JIT_CTA overrides.
int jit_cta_threads(const gsdf_program *p) {
    static const int forced = getenv("GSDF_JIT_CTA") ? atoi(getenv("GSDF_JIT_CTA")) : 0;
    if (forced >= 32 && forced <= kEvalThreads && forced % 32 == 0) return forced;
    return p->ninstr <= kSmallProgram ? kSmallCta : 256;
}

template <int P, class Gen, bool EXT>
int launch_eval_impl(const gsdf_program *p, const Gen &gen, uint64_t nwork_upper_bound, cudaStream_t st, bool pdl, uint32_t *sched,
                     unsigned long long *stamp, int threads) {
    static KernelDevCache cache;
    auto kern = k_eval<P, Gen, EXT>;
    const uint32_t smem = smem_total_bytes<P>(p->pv, threads);
    int occ = 0;
    const int rc = kernel_occupancy(cache, kern, p->device, smem, threads, &occ);
    if (rc) return rc;
    if (occ < 1) return fail(GSDF_EPROGRAM, "node program needs %u bytes of shared memory per CTA; does not fit", smem);
    uint64_t blocks = (nwork_upper_bound + threads - 1) / threads;
    blocks = std::min<uint64_t>(blocks, (uint64_t)p->sms * occ);
    ProgView pv = p->pv;
    pv.sched = sched ? sched : next_sched(p);
    pv.stamp = stamp;
    if (pdl) CU(launch_chain(true, kern, dim3((unsigned)blocks), dim3(threads), smem, st, pv, gen));
    else kern<<<(unsigned)blocks, threads, smem, st>>>(pv, gen);
    CU(cudaGetLastError());
    return 0;
}
// The same launch with a kernel compiled at run time for this program's structure (jit.cu): a cudaKernel_t from a
// context-independent library; attribute opt-in and occupancy are kept per (kernel, device) like the built-in kernels'.
struct JitLaunchFacts {
    std::mutex mu;
    std::map<std::pair<const void *, int>, bool> optin;                       // (kernel, device)
    std::map<std::tuple<const void *, int, uint32_t, int>, int> occ;          // (kernel, device, smem, threads)
};
template <int P, class Gen>
int launch_jit(const gsdf_program *p, cudaKernel_t kernel, const Gen &gen, uint64_t nwork_upper_bound, cudaStream_t st, bool pdl, uint32_t *sched,
               unsigned long long *stamp, int threads) {
    static JitLaunchFacts facts;
    const void *fn = reinterpret_cast<const void *>(kernel);
    const uint32_t smem = smem_total_bytes<P>(p->pv, threads);
    int occ = 0;
    {
        std::lock_guard<std::mutex> lk(facts.mu);
        bool &done = facts.optin[{fn, p->device}];
        if (!done) {
            DevInfo di;
            const int rc = device_info(p->device, &di);
            if (rc) return rc;
            CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin));
            done = true;
        }
        auto key = std::make_tuple(fn, p->device, smem, threads);
        auto it = facts.occ.find(key);
        if (it == facts.occ.end()) {
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem));
            facts.occ[key] = occ;
        } else occ = it->second;
    }
    if (occ < 1) return fail(GSDF_EPROGRAM, "node program needs %u bytes of shared memory per CTA; does not fit", smem);
    uint64_t blocks = (nwork_upper_bound + threads - 1) / threads;
    blocks = std::min<uint64_t>(blocks, (uint64_t)p->sms * occ);
    ProgView pv = p->pv;
    pv.sched = sched ? sched : next_sched(p);
    pv.stamp = stamp;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    Gen g = gen;
    void *args[2] = {&pv, &g};
    CU(cudaLaunchKernelExC(&cfg, fn, args));
    return 0;
}
inline const JitEntry *jit_of(const gsdf_program *p) { return (p->jit && p->jit->key == p->skey) ? p->jit.get() : nullptr; }

template <int P, class Gen>
int launch_eval(const gsdf_program *p, const Gen &gen, uint64_t nwork_upper_bound, cudaStream_t st, bool pdl, uint32_t *sched,
                unsigned long long *stamp = nullptr, int threads = kEvalThreads) {
    if (nwork_upper_bound == 0) return 0;
    return p->needs_ext ? launch_eval_impl<P, Gen, true>(p, gen, nwork_upper_bound, st, pdl, sched, stamp, threads)
                        : launch_eval_impl<P, Gen, false>(p, gen, nwork_upper_bound, st, pdl, sched, stamp, threads);
}

template <bool EXT>
int launch_prune_fine_impl(const gsdf_program *p, const PruneFine &g, cudaStream_t st, bool pdl, uint32_t *sched, unsigned long long *stamp) {
    static KernelDevCache cache;
    auto kern = k_prune_fine<EXT>;
    const uint32_t smem = smem_total_bytes<1>(p->pv, kEvalThreads) + prune_fine_extra_bytes(kEvalThreads);
    int occ = 0;
    const int rc = kernel_occupancy(cache, kern, p->device, smem, kEvalThreads, &occ);
    if (rc) return rc;
    if (occ < 1) return fail(GSDF_EPROGRAM, "node program needs %u bytes of shared memory per CTA; does not fit", smem);
    const uint64_t nwork = (uint64_t)g.L.nwx * 32u * g.L.ncy * g.L.ncz;
    uint64_t blocks = (nwork + kEvalThreads - 1) / kEvalThreads;
    blocks = std::max<uint64_t>(std::min<uint64_t>(blocks, (uint64_t)p->sms * occ), 1);
    ProgView pv = p->pv;
    pv.sched = sched ? sched : next_sched(p);
    pv.stamp = stamp;
    CU(launch_chain(pdl, kern, dim3((unsigned)blocks), dim3(kEvalThreads), smem, st, pv, g));
    CU(cudaGetLastError());
    return 0;
}

// Streaming Evaluate (k_eval_stream): persistent grid, one resident wave, tiles dealt round-robin.
template <int DIM, bool EXT>
int launch_stream_impl(const gsdf_program *p, const float *d_pos, float *d_dist, uint64_t n, cudaStream_t st) {
    static KernelDevCache cache;
    auto kern = k_eval_stream<DIM, EXT>;
    const uint32_t base = smem_total_bytes<4>(p->pv, kEvalThreads);
    const uint32_t smem = ((base + 127u) & ~127u) + 2u * stream_stage_bytes<DIM>(kEvalThreads);
    DevInfo di;
    int rc = device_info(p->device, &di);
    if (rc) return rc;
    if (smem > (uint32_t)di.smem_optin) return 1;  // does not fit: caller falls back to k_eval
    int occ = 0;
    rc = kernel_occupancy(cache, kern, p->device, smem, kEvalThreads, &occ);
    if (rc) return rc;
    if (occ < 1) return 1;
    const uint64_t tiles = (n + (uint64_t)kEvalThreads * 4 - 1) / ((uint64_t)kEvalThreads * 4);
    const unsigned blocks = (unsigned)std::min<uint64_t>(tiles, (uint64_t)p->sms * occ);
    ProgView pv = p->pv;
    pv.stamp = nullptr;
    kern<<<blocks, kEvalThreads, smem, st>>>(pv, d_pos, d_dist, n);
    CU(cudaGetLastError());
    return 0;
}
template <int DIM>
int launch_stream(const gsdf_program *p, const float *d_pos, float *d_dist, uint64_t n, cudaStream_t st) {
    static const bool off = getenv("GSDF_NO_STREAM") != nullptr;  // A/B switch
    if (off || ((((uintptr_t)d_pos | (uintptr_t)d_dist) & 15) != 0) || n < (uint64_t)kEvalThreads * 4) return 1;
    return p->needs_ext ? launch_stream_impl<DIM, true>(p, d_pos, d_dist, n, st) : launch_stream_impl<DIM, false>(p, d_pos, d_dist, n, st);
}

}  // namespace

uint32_t *next_sched(const gsdf_program *p, int *slot_index) {
    gsdf_program *q = const_cast<gsdf_program *>(p);
    const uint32_t slot = q->sched_next.fetch_add(1u, std::memory_order_relaxed) % (uint32_t)kSchedRing;
    if (slot_index) *slot_index = (int)slot;
    return p->d_sched + 2 * slot;
}

int launch_points3(const gsdf_program *p, const GenPoints3 &g, uint64_t nwork, cudaStream_t st, uint32_t *sched) { return launch_eval<4>(p, g, nwork, st, false, sched); }
int launch_points2(const gsdf_program *p, const GenPoints2 &g, uint64_t nwork, cudaStream_t st, uint32_t *sched) { return launch_eval<4>(p, g, nwork, st, false, sched); }
int launch_grid4(const gsdf_program *p, const GenGrid<4> &g, uint64_t nwork, cudaStream_t st, bool pdl, uint32_t *sched, unsigned long long *stamp) {
    if (nwork && jit_of(p)) return launch_jit<4>(p, jit_of(p)->grid4, g, nwork, st, pdl, sched, stamp, jit_cta_threads(p));
    return launch_eval<4>(p, g, nwork, st, pdl, sched, stamp, eval_cta_threads(p, false));
}
int launch_grid1(const gsdf_program *p, const GenGrid<1> &g, uint64_t nwork, cudaStream_t st, bool pdl, uint32_t *sched, unsigned long long *stamp) {
    if (nwork && jit_of(p)) return launch_jit<1>(p, jit_of(p)->grid1, g, nwork, st, pdl, sched, stamp, jit_cta_threads(p));
    return launch_eval<1>(p, g, nwork, st, pdl, sched, stamp, eval_cta_threads(p, true));
}
bool has_grid2(const gsdf_program *p) { return jit_of(p) && jit_of(p)->grid2; }
int launch_grid2(const gsdf_program *p, const GenGrid<2> &g, uint64_t nwork, cudaStream_t st, bool pdl, uint32_t *sched, unsigned long long *stamp) {
    if (!has_grid2(p)) return fail(GSDF_EUNSUPPORTED, "two corners per thread needs the specialised kernels");
    return nwork ? launch_jit<2>(p, jit_of(p)->grid2, g, nwork, st, pdl, sched, stamp, jit_cta_threads(p)) : 0;
}
int launch_prune_fine(const gsdf_program *p, const PruneFine &g, cudaStream_t st, bool pdl, uint32_t *sched, unsigned long long *stamp) {
    return p->needs_ext ? launch_prune_fine_impl<true>(p, g, st, pdl, sched, stamp) : launch_prune_fine_impl<false>(p, g, st, pdl, sched, stamp);
}
int eval_cta_slots(const gsdf_program *p, int *slots, int *threads_out) {
    static KernelDevCache cache;  // occupancy of the P = 4 lattice kernel (the P = 1 form is never lower)
    int occ = 0;
    const int threads = jit_of(p) ? jit_cta_threads(p) : eval_cta_threads(p, false);
    const uint32_t smem = smem_total_bytes<4>(p->pv, threads);
    const int rc = p->needs_ext ? kernel_occupancy(cache, k_eval<4, GenGrid<4>, true>, p->device, smem, threads, &occ)
                                : kernel_occupancy(cache, k_eval<4, GenGrid<4>, false>, p->device, smem, threads, &occ);
    if (rc) return rc;
    *slots = p->sms * std::max(occ, 1);
    if (threads_out) *threads_out = threads;
    return 0;
}
int launch_centers(const gsdf_program *p, const GenCenters &g, uint64_t nwork, cudaStream_t st, bool pdl, uint32_t *sched, unsigned long long *stamp) {
    if (nwork && jit_of(p)) return launch_jit<1>(p, jit_of(p)->centers, g, nwork, st, pdl, sched, stamp, jit_cta_threads(p));
    return launch_eval<1>(p, g, nwork, st, pdl, sched, stamp, eval_cta_threads(p, true));
}
int launch_image(const gsdf_program *p, const GenImage &g, uint64_t nwork, cudaStream_t st, uint32_t *sched) { return launch_eval<4>(p, g, nwork, st, false, sched); }
int launch_dc(const gsdf_program *p, const GenDC &g, uint64_t nwork, cudaStream_t st, uint32_t *sched) { return launch_eval<4>(p, g, nwork, st, false, sched); }
int launch_stream3(const gsdf_program *p, const float *d_pos, float *d_dist, uint64_t n, cudaStream_t st) { return launch_stream<3>(p, d_pos, d_dist, n, st); }
int launch_stream2(const gsdf_program *p, const float *d_pos, float *d_dist, uint64_t n, cudaStream_t st) { return launch_stream<2>(p, d_pos, d_dist, n, st); }

}  // namespace gsdfi
