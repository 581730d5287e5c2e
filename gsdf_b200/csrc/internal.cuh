// internal.cuh -- host-side declarations shared by the translation units of libgsdfb200.so:
//   capi.cu    error state, devices, programs, Evaluate / lattice / image entry points, pinned host memory
//   eval.cu    the interpreter kernels (slow to compile: every generator x {default, EXT}) and their launchers
//   mesher.cu  marching-cubes / scan / STL / dual-contouring kernels, the mesher and the multi-device mesher
// Nothing here is part of the C ABI (include/gsdf_b200.h is).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gsdf_b200.h"
#include "../../include/gsdf_program.h"
#include "dc_gen.cuh"
#include "generators.cuh"

namespace gsdfi {

int fail(int code, const char *fmt, ...);  // sets the calling thread's gsdf_last_error() text, returns code
#define CU(call)                                                                                                \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess) return gsdfi::fail(GSDF_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_));  \
    } while (0)

struct DevInfo {
    int sms = 0;         // multiprocessors
    int smem_optin = 0;  // cudaDevAttrMaxSharedMemoryPerBlockOptin
};
// Per-device facts, queried once per device under a lock (function attributes and SM counts are per device: a process
// that drives several GPUs from several threads must never reuse another device's numbers).
int device_info(int dev, DevInfo *out);
int default_device();  // the calling thread's default device (gsdf_set_device), 0 if never set

// Every compute entry point starts here: select the handle's device and drop any stale NON-sticky error another library
// (or a teardown path) left in this thread's runtime state, so that the cudaGetLastError() checks behind our launches
// report our launches only. Sticky errors (a faulted context) are not cleared by this and still surface.
inline cudaError_t use_device(int dev) {
    const cudaError_t e = cudaSetDevice(dev);
    if (e == cudaSuccess) (void)cudaGetLastError();
    return e;
}
int ensure_device(int dev);  // checks that a CUDA device exists (no CPU fallback), selects dev

template <class T>
int grow(T *&ptr, size_t &cap, size_t need) {
    if (need <= cap) return 0;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = need + need / 8;
    cudaError_t e = cudaMalloc((void **)&ptr, want * sizeof(T));
    if (e != cudaSuccess) return fail(GSDF_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
    cap = want;
    return 0;
}

inline unsigned grid_for(int sms, uint64_t items, int per_block, int waves = 8) {
    uint64_t b = (items + per_block - 1) / per_block;
    b = std::min<uint64_t>(b, (uint64_t)sms * waves);
    return (unsigned)std::max<uint64_t>(b, 1);
}

// Launch with (pdl) or without the programmatic-stream-serialization attribute: with it the kernel may become resident
// while its predecessor on the stream drains and runs up to its pdl_wait() (generators.cuh); captured into a CUDA graph
// the attribute becomes a programmatic dependency edge.
template <class... KArgs, class... Args>
cudaError_t launch_chain(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

inline gsdfk::Lat make_lat(const gsdf_lattice *lat, int k0, int k1, int pitch, bool vec) {
    gsdfk::Lat L;
    L.ox = lat->origin[0]; L.oy = lat->origin[1]; L.oz = lat->origin[2]; L.res = lat->res;
    L.nx = lat->n[0]; L.ny = lat->n[1]; L.nz = lat->n[2];
    L.k0 = k0; L.nk = k1 - k0;
    L.nqx = (lat->n[0] + 1 + 3) / 4;
    L.pitch = pitch;
    L.vec = vec ? 1 : 0;
    L.hq = 0;
    L.fdiv = (uint64_t)L.nqx * (uint64_t)(L.ny + 1) * (uint64_t)(L.nk > 0 ? L.nk : 1) < (1ull << 31) ? 1u : 0u;
    gsdfk::fastdiv_init((uint32_t)L.nqx, L.nqx_mul, L.nqx_shr);
    gsdfk::fastdiv_init((uint32_t)(L.ny + 1), L.nyp_mul, L.nyp_shr);
    return L;
}

}  // namespace gsdfi

namespace gsdfi {
// Kernels compiled at run time for one program structure (jit.cu); shared by every program and device with that structure.
constexpr int kJitKernels = 4;
struct JitEntry {
    std::string key;
    cudaLibrary_t lib[kJitKernels] = {};
    cudaKernel_t grid4 = nullptr, grid1 = nullptr, centers = nullptr, grid2 = nullptr;
    ~JitEntry();
};
}  // namespace gsdfi

constexpr int kSchedRing = 64;  // scheduler slots of a program (one pair of counters per in-flight interpreter launch)

struct gsdf_program {
    int device = 0;
    int sms = 0;                  // multiprocessors of `device`
    uint8_t *d_blob = nullptr;
    gsdfk::ProgView pv{};         // pv.sched is filled per launch (launch_* below)
    int dim = 3;
    uint32_t ninstr = 0;
    std::atomic<uint64_t> evals{0};
    cudaStream_t stream = nullptr;
    float *d_pos = nullptr, *d_dist = nullptr;
    size_t pos_cap = 0, dist_cap = 0;
    // Work-tile schedulers of k_eval: kSchedRing self-resetting counter pairs. Every interpreter launch that does not bring
    // its own pair (the mesher does) takes the next slot, so launches of one program that are in flight on different
    // streams never share a counter; a slot is reused after kSchedRing further launches.
    uint32_t *d_sched = nullptr;
    std::atomic<uint32_t> sched_next{0};
    size_t blob_cap = 0;          // bytes allocated at d_blob
    bool needs_ext = false;       // program contains ellipse2D / quadbezier2d -> EXT interpreter instantiation
    // run-time specialisation (jit.cu): host copy of the instruction chunks, the structural key they give, and the compiled
    // kernels in use (valid while jit->key == skey; an update that changes the structure drops them)
    std::vector<uint32_t> h_words;
    std::string skey;
    std::shared_ptr<gsdfi::JitEntry> jit;
    // Streams other than `stream` that may still be reading d_blob: caller-supplied streams of the *_device entry points
    // (one event per scheduler slot, recorded behind the launch) and the streams of meshers created on this program
    // (they register their completion event). gsdf_program_update / destroy wait for all of them.
    cudaEvent_t user_ev[kSchedRing] = {};
    std::atomic<uint64_t> user_dirty{0};
    std::mutex dep_mu;
    struct Dependent { cudaEvent_t ev; gsdf_program **ref; };  // ref: the dependent's pointer to this program, nulled if the program goes first
    std::vector<Dependent> deps;
    // asynchronous upload (program_update_async): pinned staging of the blob and the event the readers' streams wait for
    uint8_t *h_blob = nullptr;
    size_t h_blob_cap = 0;
    cudaEvent_t upload_ev = nullptr;
    bool upload_ev_recorded = false;
    // pinned staging of the pipelined host Evaluate (capi.cu)
    struct EvalSlot {
        float *h_pos = nullptr, *h_dist = nullptr, *d_pos = nullptr, *d_dist = nullptr;
        size_t cap = 0;  // points
        cudaStream_t st = nullptr;
        cudaEvent_t done = nullptr;
    } slot[4][3];   // [lane][slot]: up to four host threads, three chunks in flight each
};

namespace gsdfi {

void program_add_dependent(gsdf_program *p, cudaEvent_t ev, gsdf_program **ref);
void program_remove_dependent(gsdf_program *p, cudaEvent_t ev);
// gsdf_program_update without a host synchronisation when the new program has the layout of the old one (same sizes, stack
// slots and interpreter): the bytes go through the handle's pinned staging onto its stream behind a device-side wait for
// every registered reader, and upload_ev is recorded for the next launches to wait on. Anything else: the synchronous path.
int program_update_async(gsdf_program *p, const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats);
int check_program_blob(const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats, int dim);
// jit.cu
std::string program_structure_key(const gsdf_program_header &h, const uint32_t *chunks, bool ext, bool stage_aux);
int program_specialize(gsdf_program *p);
// keeps h_words / skey of a program current and drops a specialisation the new structure no longer matches
void program_note_structure(gsdf_program *p, const gsdf_program_header &h, const uint32_t *chunks);
// Waits until nothing on any stream can still be reading the program's device buffers.
int program_quiesce(gsdf_program *p);

// ---- interpreter launchers (eval.cu). sched: counter pair owned by the caller, or nullptr = next ring slot of the program.
// nwork_upper_bound sizes the persistent grid (at most one resident wave). Returns 0 or a gsdf_status.
int launch_points3(const gsdf_program *p, const gsdfk::GenPoints3 &g, uint64_t nwork, cudaStream_t st, uint32_t *sched);
int launch_points2(const gsdf_program *p, const gsdfk::GenPoints2 &g, uint64_t nwork, cudaStream_t st, uint32_t *sched);
int launch_grid4(const gsdf_program *p, const gsdfk::GenGrid<4> &g, uint64_t nwork, cudaStream_t st, bool pdl, uint32_t *sched,
                 unsigned long long *stamp = nullptr);
// the same lattice evaluation with ONE corner per thread (4 work items per quad): for listed work that fills less than about
// one resident wave, where the render is bound by the latency of a tile, not by throughput. nwork counts quads x 4.
// two corners per thread: run-time compiled kernels only (has_grid2)
bool has_grid2(const gsdf_program *p);
int launch_grid2(const gsdf_program *p, const gsdfk::GenGrid<2> &g, uint64_t nwork, cudaStream_t st, bool pdl, uint32_t *sched, unsigned long long *stamp);
int launch_grid1(const gsdf_program *p, const gsdfk::GenGrid<1> &g, uint64_t nwork, cudaStream_t st, bool pdl, uint32_t *sched,
                 unsigned long long *stamp = nullptr);
// resident CTA slots of the lattice-evaluation kernel for this program (SMs x occupancy) and the CTA size it is launched with
int eval_cta_slots(const gsdf_program *p, int *slots, int *threads = nullptr);
int launch_centers(const gsdf_program *p, const gsdfk::GenCenters &g, uint64_t nwork, cudaStream_t st, bool pdl, uint32_t *sched,
                   unsigned long long *stamp = nullptr);
// prune levels 3 and 2 of a plan that ends with level 2, in one launch (k_prune_fine)
int launch_prune_fine(const gsdf_program *p, const gsdfk::PruneFine &g, cudaStream_t st, bool pdl, uint32_t *sched, unsigned long long *stamp = nullptr);
int launch_image(const gsdf_program *p, const gsdfk::GenImage &g, uint64_t nwork, cudaStream_t st, uint32_t *sched);
int launch_dc(const gsdf_program *p, const gsdfk::GenDC &g, uint64_t nwork, cudaStream_t st, uint32_t *sched);
// streaming Evaluate: 0 launched, 1 not applicable (caller uses launch_points*), <0 error
int launch_stream3(const gsdf_program *p, const float *d_pos, float *d_dist, uint64_t n, cudaStream_t st);
int launch_stream2(const gsdf_program *p, const float *d_pos, float *d_dist, uint64_t n, cudaStream_t st);
// the next scheduler slot of the program's ring (device pointer to its counter pair) and its index
uint32_t *next_sched(const gsdf_program *p, int *slot_index = nullptr);

}  // namespace gsdfi
