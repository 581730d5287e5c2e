// capi.cu -- implementation of the C ABI declared in include/gsdf_b200.h (device layer).
// Handles own device buffers (grow-never-shrink) and one stream each; no CPU fallback exists: every compute entry
// point fails with GSDF_ECUDA when no CUDA device is usable.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gsdf_b200.h"
#include "../../include/gsdf_program.h"
#include "kernels.cuh"
#include "dualcontour.cuh"

using namespace gsdfk;

namespace {

thread_local std::string g_err;
int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(GSDF_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_));   \
    } while (0)

int g_device = 0;
int g_sms = 0;

// Every compute entry point starts here: select the handle's device and drop any stale NON-sticky error another library
// (or a teardown path) left in this thread's runtime state, so that the cudaGetLastError() checks behind our launches
// report our launches only. Sticky errors (a faulted context) are not cleared by this and still surface.
cudaError_t use_device(int dev) {
    const cudaError_t e = cudaSetDevice(dev);
    if (e == cudaSuccess) (void)cudaGetLastError();
    return e;
}

int ensure_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return fail(GSDF_ECUDA, "no CUDA device available (%s); libgsdfb200 has no CPU fallback", cudaGetErrorString(e));
    CU(cudaSetDevice(g_device));
    if (!g_sms) {
        cudaDeviceProp p;
        CU(cudaGetDeviceProperties(&p, g_device));
        g_sms = p.multiProcessorCount;
    }
    return 0;
}

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

}  // namespace

struct gsdf_program {
    int device = 0;
    uint8_t *d_blob = nullptr;
    ProgView pv{};
    int dim = 3;
    uint32_t ninstr = 0;
    uint64_t evals = 0;
    cudaStream_t stream = nullptr;
    float *d_pos = nullptr, *d_dist = nullptr;
    size_t pos_cap = 0, dist_cap = 0;
    uint32_t *d_sched = nullptr;  // work-tile scheduler of k_eval (self-resetting)
    size_t blob_cap = 0;          // bytes allocated at d_blob
    bool needs_ext = false;       // program contains ellipse2D / quadbezier2d -> EXT interpreter instantiation
};

namespace {

// persistent launch: at most one resident wave of CTAs; they pull 256-item tiles from the program's scheduler
// Launch with (pdl) or without the programmatic-stream-serialization attribute: with it the kernel may become resident
// while its predecessor on the stream drains and runs up to its pdl_wait() (kernels.cuh); captured into a CUDA graph the
// attribute becomes a programmatic dependency edge.
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

template <int P, class Gen, bool EXT>
int launch_eval_impl(const gsdf_program *p, const Gen &gen, uint64_t nwork_upper_bound, cudaStream_t st, bool pdl) {
    auto kern = k_eval<P, Gen, EXT>;
    const uint32_t smem = smem_total_bytes<P>(p->pv, kEvalThreads);
    static thread_local uint32_t cached_smem = 0xffffffffu;
    static thread_local int cached_occ = 0;
    if (cached_smem != smem) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<uint32_t>(smem, 48 * 1024)));
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kEvalThreads, smem));
        if (occ < 1) return fail(GSDF_EPROGRAM, "node program needs %u bytes of shared memory per CTA; does not fit", smem);
        cached_occ = occ;
        cached_smem = smem;
    }
    uint64_t blocks = (nwork_upper_bound + kEvalThreads - 1) / kEvalThreads;
    blocks = std::min<uint64_t>(blocks, (uint64_t)g_sms * cached_occ);
    if (pdl) CU(launch_chain(true, kern, dim3((unsigned)blocks), dim3(kEvalThreads), smem, st, p->pv, gen));
    else kern<<<(unsigned)blocks, kEvalThreads, smem, st>>>(p->pv, gen);
    CU(cudaGetLastError());
    return 0;
}
// persistent launch: at most one resident wave of CTAs; they pull tiles from the program's scheduler
template <int P, class Gen>
int launch_eval(const gsdf_program *p, const Gen &gen, uint64_t nwork_upper_bound, cudaStream_t st, bool pdl = false) {
    if (nwork_upper_bound == 0) return 0;
    return p->needs_ext ? launch_eval_impl<P, Gen, true>(p, gen, nwork_upper_bound, st, pdl)
                        : launch_eval_impl<P, Gen, false>(p, gen, nwork_upper_bound, st, pdl);
}

// Streaming Evaluate (k_eval_stream): persistent grid, one resident wave, tiles dealt round-robin.
template <int DIM, bool EXT>
int launch_stream_impl(const gsdf_program *p, const float *d_pos, float *d_dist, uint64_t n, cudaStream_t st) {
    auto kern = k_eval_stream<DIM, EXT>;
    const uint32_t base = smem_total_bytes<4>(p->pv, kEvalThreads);
    const uint32_t smem = ((base + 127u) & ~127u) + 2u * stream_stage_bytes<DIM>(kEvalThreads);
    static thread_local uint32_t cached_smem = 0xffffffffu;
    static thread_local int cached_occ = 0;
    if (cached_smem != smem) {
        if (smem > 227u * 1024u) { cached_smem = 0xffffffffu; return 1; }  // does not fit: caller falls back to k_eval
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<uint32_t>(smem, 48 * 1024)));
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kEvalThreads, smem));
        if (occ < 1) return 1;
        cached_occ = occ;
        cached_smem = smem;
    }
    const uint64_t tiles = (n + (uint64_t)kEvalThreads * 4 - 1) / ((uint64_t)kEvalThreads * 4);
    const unsigned blocks = (unsigned)std::min<uint64_t>(tiles, (uint64_t)g_sms * cached_occ);
    kern<<<blocks, kEvalThreads, smem, st>>>(p->pv, d_pos, d_dist, n);
    CU(cudaGetLastError());
    return 0;
}
// returns 0 launched, 1 not applicable (caller uses the generic kernel), <0 error
template <int DIM>
int launch_stream(const gsdf_program *p, const float *d_pos, float *d_dist, uint64_t n, cudaStream_t st) {
    static const bool off = getenv("GSDF_NO_STREAM") != nullptr;  // A/B switch
    if (off || ((((uintptr_t)d_pos | (uintptr_t)d_dist) & 15) != 0) || n < (uint64_t)kEvalThreads * 4) return 1;
    return p->needs_ext ? launch_stream_impl<DIM, true>(p, d_pos, d_dist, n, st) : launch_stream_impl<DIM, false>(p, d_pos, d_dist, n, st);
}

int validate_program(const gsdf_program_header &h, const uint32_t *chunks, size_t aux_floats) {
    uint32_t pc = 0, n = 0;
    bool ended = false;
    std::vector<uint8_t> starts(h.nchunks, 0);  // instruction boundaries, for slab-guard jump targets
    for (uint32_t q = 0; q < h.nchunks;) {
        starts[q] = 1;
        const uint32_t len = (chunks[4 * q] >> 8) & 0xff;
        if (len < 1) break;
        q += len;
    }
    while (pc < h.nchunks) {
        const uint32_t w0 = chunks[4 * pc], op = w0 & 0xff, len = (w0 >> 8) & 0xff;
        if (op >= GSDF_OP__COUNT) return fail(GSDF_EPROGRAM, "instruction %u: unknown opcode %u", n, op);
        if (len < 1 || pc + len > h.nchunks) return fail(GSDF_EPROGRAM, "instruction %u: bad length %u", n, len);
        static const uint8_t need2[] = {GSDF_OP_BOX, GSDF_OP_BOXFRAME, GSDF_OP_CYLINDER, GSDF_OP_HEX, GSDF_OP_DIAMOND2D, GSDF_OP_TRANSLATE,
                                        GSDF_OP_ROTATE2D, GSDF_OP_ELONGATE, GSDF_OP_ARRAY2D_VAR, GSDF_OP_CIRC_ENTER, GSDF_OP_SCREW_ENTER, GSDF_OP_BBOX_GUARD2D};
        static const uint8_t need3[] = {GSDF_OP_LINE2D, GSDF_OP_ARC2D, GSDF_OP_ARRAY_VAR};
        uint32_t want = 1;
        for (uint8_t o : need2) if (o == op) want = 2;
        for (uint8_t o : need3) if (o == op) want = 3;
        if (op == GSDF_OP_TRANSFORM || op == GSDF_OP_BEZIERQ2D) want = 4;
        if (len != want) return fail(GSDF_EPROGRAM, "instruction %u (opcode %u): length %u, expected %u", n, op, len, want);
        if (op == GSDF_OP_POLY2D) {
            const uint64_t off = chunks[4 * pc + 1], nv = chunks[4 * pc + 2];
            if ((off & 3) || nv < 3 || off + nv * GSDF_POLY_EDGE_FLOATS > aux_floats) return fail(GSDF_EPROGRAM, "instruction %u: polygon aux range out of bounds", n);
        }
        if (op == GSDF_OP_CULL_UB2D) {
            const uint64_t off = chunks[4 * pc + 1], np = chunks[4 * pc + 2];
            if ((off & 3) || np < 2 || (np & 1) || off + np * 2 > aux_floats) return fail(GSDF_EPROGRAM, "instruction %u: anchor aux range out of bounds", n);
        }
        if (op == GSDF_OP_BBOX_GUARD2D) {
            const uint32_t kind = chunks[4 * pc + 1] & 0xff, target = chunks[4 * pc + 1] >> 8;
            if (kind != GSDF_GUARD_DIFF && kind != GSDF_GUARD_MIN) return fail(GSDF_EPROGRAM, "instruction %u: unknown box guard %u", n, kind);
            if (target <= pc + len || target >= h.nchunks || !starts[target]) return fail(GSDF_EPROGRAM, "instruction %u: box guard target %u is not a later instruction", n, target);
            const uint32_t top = chunks[4 * target] & 0xff;
            if (top != (kind == GSDF_GUARD_MIN ? (uint32_t)GSDF_OP_MIN : (uint32_t)GSDF_OP_DIFF)) return fail(GSDF_EPROGRAM, "instruction %u: box guard target is not its combiner", n);
        }
        if (op == GSDF_OP_EXTRUDE_ENTER || op == GSDF_OP_SCREW_ENTER) {
            const uint32_t kind = chunks[4 * pc + 1] & 0xff, target = chunks[4 * pc + 1] >> 8;
            if (kind > GSDF_GUARD_SMOOTH_UNION) return fail(GSDF_EPROGRAM, "instruction %u: unknown slab guard %u", n, kind);
            if (kind != GSDF_GUARD_NONE && (target <= pc + len || target >= h.nchunks || !starts[target]))
                return fail(GSDF_EPROGRAM, "instruction %u: slab guard target %u is not a later instruction", n, target);
        }
        if (op == GSDF_OP_CYLINDER || op == GSDF_OP_TORUS || op == GSDF_OP_CIRCLE2D || op == GSDF_OP_SCREW_ENTER) {
            // radius reuse (gsdf_program.h, experimental): flag word is w1, for SCREW_ENTER w2
            const uint32_t fl = chunks[4 * pc + (op == GSDF_OP_SCREW_ENTER ? 2 : 1)] & (GSDF_RXY_READ | GSDF_RXY_WRITE);
#ifdef GSDF_RXY
            if (fl == (GSDF_RXY_READ | GSDF_RXY_WRITE)) return fail(GSDF_EPROGRAM, "instruction %u: radius flags READ and WRITE are exclusive", n);
#else
            if (fl) return fail(GSDF_EPROGRAM, "instruction %u: radius-reuse flags need a library built with -DGSDF_RXY (unset GSDF_RXY in the flattener's environment)", n);
#endif
        }
        if (op == GSDF_OP_LINES2D) {
            const uint64_t off = chunks[4 * pc + 1], ns = chunks[4 * pc + 2];
            if ((off & 3) || off + ns * 4 > aux_floats) return fail(GSDF_EPROGRAM, "instruction %u: lines aux range out of bounds", n);
        }
        pc += len;
        n++;
        if (op == GSDF_OP_END) { ended = true; break; }
    }
    if (!ended || pc != h.nchunks) return fail(GSDF_EPROGRAM, "program does not end with END at its last chunk");
    // Stack discipline against the header. The stream is straight-line (a firing guard skips a region whose net effect
    // on both stacks is zero), so one walk gives the depth at every instruction. The kernels size their shared-memory
    // stacks from the header: a program that pushes deeper than it declares would write outside them.
    int d = 0, dmax = 0, ps = 0, pmax = 0;
    std::vector<int> dAt(h.nchunks, -1), pAt(h.nchunks, -1);  // depths BEFORE the instruction that starts at a chunk
    n = 0;
    for (pc = 0; pc < h.nchunks; n++) {
        const uint32_t op = chunks[4 * pc] & 0xff, len = (chunks[4 * pc] >> 8) & 0xff;
        dAt[pc] = d; pAt[pc] = ps;
        int dd = 0, dp = 0, needd = 0, needp = 0;
        switch (op) {
        case GSDF_OP_SPHERE: case GSDF_OP_BOX: case GSDF_OP_BOXFRAME: case GSDF_OP_TORUS: case GSDF_OP_CYLINDER: case GSDF_OP_HEX:
        case GSDF_OP_CIRCLE2D: case GSDF_OP_RECT2D: case GSDF_OP_LINE2D: case GSDF_OP_LINES2D: case GSDF_OP_ARC2D: case GSDF_OP_EQTRI2D:
        case GSDF_OP_HEX2D: case GSDF_OP_OCT2D: case GSDF_OP_DIAMOND2D: case GSDF_OP_ROUNDX2D: case GSDF_OP_POLY2D: case GSDF_OP_ELLIPSE2D:
        case GSDF_OP_BEZIERQ2D: case GSDF_OP_CULL_UB2D: case GSDF_OP_ELONGATE: case GSDF_OP_ELONGATE2D: case GSDF_OP_EXTRUDE_ENTER:
        case GSDF_OP_SCREW_ENTER:
            dd = 1; break;
        case GSDF_OP_MIN: case GSDF_OP_MAX: case GSDF_OP_DIFF: case GSDF_OP_XOR: case GSDF_OP_SMOOTH_UNION: case GSDF_OP_SMOOTH_DIFF:
        case GSDF_OP_SMOOTH_INTERSECT: case GSDF_OP_ADD_BELOW: case GSDF_OP_EXTRUDE_EXIT: case GSDF_OP_MAX_BELOW:
            dd = -1; needd = 2; break;
        case GSDF_OP_OFFSET: case GSDF_OP_ANNULUS: case GSDF_OP_MULDIST: case GSDF_OP_SHELL_EXIT: case GSDF_OP_BBOX_GUARD2D:
            needd = 1; break;
        case GSDF_OP_PUSH_POS: case GSDF_OP_CIRC_ENTER: dp = 1; break;
        case GSDF_OP_POP_POS: dp = -1; needp = 1; break;
        case GSDF_OP_PEEK_POS: needp = 1; break;
        default: break;
        }
        if (d < needd) return fail(GSDF_EPROGRAM, "instruction %u (opcode %u): distance stack underflow", n, op);
        if (ps < needp) return fail(GSDF_EPROGRAM, "instruction %u (opcode %u): position stack underflow", n, op);
        d += dd; ps += dp;
        dmax = std::max(dmax, d); pmax = std::max(pmax, ps);
        if (op == GSDF_OP_END) break;
        pc += len;
    }
    if (d != 1 || ps != 0) return fail(GSDF_EPROGRAM, "program leaves %d distances and %d positions on its stacks (expected 1 and 0)", d, ps);
    // A firing guard jumps to its target with the `skip` flag set: the skipped region must have produced exactly the one
    // value its combiner would have consumed, and what lies between the target and that combiner may only restore p.
    n = 0;
    for (pc = 0; pc < h.nchunks; n++) {
        const uint32_t op = chunks[4 * pc] & 0xff, len = (chunks[4 * pc] >> 8) & 0xff;
        const bool slab = (op == GSDF_OP_EXTRUDE_ENTER || op == GSDF_OP_SCREW_ENTER) && (chunks[4 * pc + 1] & 0xff) != GSDF_GUARD_NONE;
        if (slab || op == GSDF_OP_BBOX_GUARD2D) {
            const uint32_t kind = chunks[4 * pc + 1] & 0xff, target = chunks[4 * pc + 1] >> 8;
            if (dAt[target] != dAt[pc] + 1 || pAt[target] != pAt[pc] || dAt[pc] < 1)
                return fail(GSDF_EPROGRAM, "instruction %u: the region its guard skips does not leave exactly one value for the combiner", n);
            uint32_t q = target;
            while (q < h.nchunks && (chunks[4 * q] & 0xff) == GSDF_OP_POP_POS) q += (chunks[4 * q] >> 8) & 0xff;
            const uint32_t comb = q < h.nchunks ? (chunks[4 * q] & 0xff) : (uint32_t)GSDF_OP_END;
            const uint32_t want = kind == GSDF_GUARD_MIN ? (uint32_t)GSDF_OP_MIN : kind == GSDF_GUARD_DIFF ? (uint32_t)GSDF_OP_DIFF : (uint32_t)GSDF_OP_SMOOTH_UNION;
            if (comb != want) return fail(GSDF_EPROGRAM, "instruction %u: guard kind %u does not lead to its combiner", n, kind);
        }
        if (op == GSDF_OP_END) break;
        pc += len;
    }
    // the top of the distance stack lives in registers and slot 0 absorbs the first push: dmax values need dmax - 1 slots
    if ((uint32_t)std::max(dmax - 1, 1) > h.dstack || (uint32_t)pmax > h.pstack)
        return fail(GSDF_EPROGRAM, "program needs %d distance and %d position stack slots, its header declares %u and %u", std::max(dmax - 1, 1), pmax, h.dstack, h.pstack);
    return 0;
}

}  // namespace

extern "C" {

#ifdef GSDF_RXY
const char *gsdf_version(void) { return "gsdf-b200 0.1 (sm_100a) +rxy"; }  // experimental radius-reuse build (gsdf_program.h)
#else
const char *gsdf_version(void) { return "gsdf-b200 0.1 (sm_100a)"; }
#endif
const char *gsdf_last_error(void) { return g_err.c_str(); }

int gsdf_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(GSDF_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

int gsdf_set_device(int device) {
    int n = gsdf_device_count();
    if (n < 0) return n;
    if (device < 0 || device >= n) return fail(GSDF_EINVAL, "device %d out of range (have %d)", device, n);
    g_device = device;
    g_sms = 0;
    return ensure_device();
}

static int parse_blob(const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats, gsdf_program_header &h, const uint32_t *&chunks) {
    if (!blob || blob_bytes < sizeof(gsdf_program_header)) return fail(GSDF_EINVAL, "program blob is NULL or too short");
    std::memcpy(&h, blob, sizeof h);
    if (h.magic != GSDF_PROGRAM_MAGIC || h.version != GSDF_PROGRAM_VERSION) return fail(GSDF_EPROGRAM, "bad program magic/version");
    if (h.nchunks == 0 || blob_bytes != sizeof h + (size_t)h.nchunks * 16) return fail(GSDF_EPROGRAM, "program size mismatch");
    if (h.dim != 2 && h.dim != 3) return fail(GSDF_EPROGRAM, "program dim must be 2 or 3");
    if (aux_floats && !aux) return fail(GSDF_EINVAL, "aux is NULL");
    if (aux_floats & 3) return fail(GSDF_EPROGRAM, "aux length must be a multiple of 4 floats");
    if (h.dstack < 1 || h.dstack > 64 || h.pstack > 32) return fail(GSDF_EPROGRAM, "stack depth out of range (d=%u p=%u)", h.dstack, h.pstack);
    chunks = reinterpret_cast<const uint32_t *>(static_cast<const uint8_t *>(blob) + sizeof h);
    const size_t prog_bytes = (size_t)h.nchunks * 16;
    const uint32_t stacks = kEvalThreads * 4u * 4u * (h.dstack + 3u * h.pstack + kRxySlots);
    if (prog_bytes + stacks + 16 > 200 * 1024) return fail(GSDF_EPROGRAM, "program too large for shared memory");
    return validate_program(h, chunks, aux_floats);
}

// copies chunks + aux into p->d_blob (growing it if needed) and refreshes the kernel-side view
static int upload_blob(gsdf_program *p, const gsdf_program_header &h, const uint32_t *chunks, const float *aux, size_t aux_floats) {
    const size_t prog_bytes = (size_t)h.nchunks * 16, aux_bytes = aux_floats * 4;
    if (prog_bytes + aux_bytes + 16 > p->blob_cap) {
        if (p->d_blob) cudaFree(p->d_blob);
        p->d_blob = nullptr;
        p->blob_cap = 0;
        const size_t want = std::max<size_t>(4096, 2 * (prog_bytes + aux_bytes + 16));
        cudaError_t e = cudaMalloc((void **)&p->d_blob, want);
        if (e != cudaSuccess) return fail(GSDF_ENOMEM, "cudaMalloc program: %s", cudaGetErrorString(e));
        p->blob_cap = want;
    }
    CU(cudaMemcpyAsync(p->d_blob, chunks, prog_bytes, cudaMemcpyHostToDevice, p->stream));
    if (aux_bytes) CU(cudaMemcpyAsync(p->d_blob + prog_bytes, aux, aux_bytes, cudaMemcpyHostToDevice, p->stream));
    CU(cudaStreamSynchronize(p->stream));  // the caller's buffers may go away after we return
    p->dim = (int)h.dim;
    p->ninstr = h.ninstr;
    p->needs_ext = false;
    for (uint32_t pc = 0; pc < h.nchunks;) {
        const uint32_t op = chunks[4 * pc] & 0xff, len = (chunks[4 * pc] >> 8) & 0xff;
        if (op == GSDF_OP_ELLIPSE2D || op == GSDF_OP_BEZIERQ2D) p->needs_ext = true;
        pc += len ? len : 1;
    }
    p->pv.g_prog = reinterpret_cast<const uint4 *>(p->d_blob);
    p->pv.prog_bytes = (uint32_t)prog_bytes;
    p->pv.aux_bytes = (uint32_t)aux_bytes;
    p->pv.dslots = h.dstack;
    p->pv.pslots = h.pstack;
    // stage aux with the program when program + aux + stacks stay under ~100 KB (>= 2 CTAs/SM)
    const uint32_t stacks = kEvalThreads * 4u * 4u * (h.dstack + 3u * h.pstack + kRxySlots);
    p->pv.stage_aux = (prog_bytes + aux_bytes + stacks + 16 <= 100 * 1024) ? 1u : 0u;
    return 0;
}

int gsdf_program_create(const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats, gsdf_program **out) {
    if (!out) return fail(GSDF_EINVAL, "gsdf_program_create: out is NULL");
    gsdf_program_header h;
    const uint32_t *chunks = nullptr;
    int rc = parse_blob(blob, blob_bytes, aux, aux_floats, h, chunks);
    if (rc) return rc;
    rc = ensure_device();
    if (rc) return rc;
    gsdf_program *p = new gsdf_program();
    p->device = g_device;
    cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_sched, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(p->d_sched, 0, 2 * sizeof(uint32_t));
    if (e != cudaSuccess) { gsdf_program_destroy(p); return fail(GSDF_ECUDA, "program setup: %s", cudaGetErrorString(e)); }
    p->pv.sched = p->d_sched;
    rc = upload_blob(p, h, chunks, aux, aux_floats);
    if (rc) { gsdf_program_destroy(p); return rc; }
    *out = p;
    return 0;
}

int gsdf_program_update(gsdf_program *p, const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats) {
    if (!p) return fail(GSDF_EINVAL, "gsdf_program_update: NULL program");
    gsdf_program_header h;
    const uint32_t *chunks = nullptr;
    int rc = parse_blob(blob, blob_bytes, aux, aux_floats, h, chunks);
    if (rc) return rc;
    if ((int)h.dim != p->dim) return fail(GSDF_EINVAL, "cannot change a %dD program into a %dD one", p->dim, (int)h.dim);
    CU(use_device(p->device));
    CU(cudaStreamSynchronize(p->stream));  // nothing may still be reading the old program
    return upload_blob(p, h, chunks, aux, aux_floats);
}

void gsdf_program_destroy(gsdf_program *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamDestroy(p->stream);
    cudaFree(p->d_blob);
    cudaFree(p->d_sched);
    cudaFree(p->d_pos);
    cudaFree(p->d_dist);
    delete p;
    (void)cudaGetLastError();
}

uint64_t gsdf_program_evaluations(const gsdf_program *p) { return p ? p->evals : 0; }

int gsdf_eval3_device(gsdf_program *p, const float *d_pos, float *d_dist, size_t n, void *stream) {
    if (!p || !d_pos || !d_dist) return fail(GSDF_EINVAL, "gsdf_eval3_device: NULL argument");
    if (p->dim != 3) return fail(GSDF_EINVAL, "program is not 3D");
    if (n == 0) return fail(GSDF_EEMPTY, "empty buffers");
    CU(use_device(p->device));
    const int src = launch_stream<3>(p, d_pos, d_dist, (uint64_t)n, stream ? (cudaStream_t)stream : p->stream);
    if (src <= 0) return src;
    GenPoints3 g{d_pos, d_dist, (uint64_t)n, (((uintptr_t)d_pos | (uintptr_t)d_dist) & 15) == 0 ? 1 : 0};
    return launch_eval<4>(p, g, (n + 3) / 4, stream ? (cudaStream_t)stream : p->stream);
}

int gsdf_eval2_device(gsdf_program *p, const float *d_pos, float *d_dist, size_t n, void *stream) {
    if (!p || !d_pos || !d_dist) return fail(GSDF_EINVAL, "gsdf_eval2_device: NULL argument");
    if (p->dim != 2) return fail(GSDF_EINVAL, "program is not 2D");
    if (n == 0) return fail(GSDF_EEMPTY, "empty buffers");
    CU(use_device(p->device));
    const int src = launch_stream<2>(p, d_pos, d_dist, (uint64_t)n, stream ? (cudaStream_t)stream : p->stream);
    if (src <= 0) return src;
    GenPoints2 g{d_pos, d_dist, (uint64_t)n, (((uintptr_t)d_pos | (uintptr_t)d_dist) & 15) == 0 ? 1 : 0};
    return launch_eval<4>(p, g, (n + 3) / 4, stream ? (cudaStream_t)stream : p->stream);
}

static int eval_host(gsdf_program *p, const float *pos, float *dist, size_t n, int dim) {
    if (!p || !pos || !dist) return fail(GSDF_EINVAL, "gsdf_eval: NULL argument");
    if (p->dim != dim) return fail(GSDF_EINVAL, "program is %dD, called as %dD", p->dim, dim);
    if (n == 0) return fail(GSDF_EEMPTY, "empty buffers");
    CU(use_device(p->device));
    int rc = grow(p->d_pos, p->pos_cap, n * 3);
    if (rc) return rc;
    rc = grow(p->d_dist, p->dist_cap, n);
    if (rc) return rc;
    CU(cudaMemcpyAsync(p->d_pos, pos, n * dim * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    rc = dim == 3 ? gsdf_eval3_device(p, p->d_pos, p->d_dist, n, p->stream) : gsdf_eval2_device(p, p->d_pos, p->d_dist, n, p->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(dist, p->d_dist, n * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    p->evals += n;
    return 0;
}
int gsdf_eval3(gsdf_program *p, const float *pos, float *dist, size_t n) { return eval_host(p, pos, dist, n, 3); }
int gsdf_eval2(gsdf_program *p, const float *pos, float *dist, size_t n) { return eval_host(p, pos, dist, n, 2); }

// ------------------------------------------------------------------------------------------------ lattice
int gsdf_lattice_from_bounds(const float bbmin[3], const float bbmax[3], float res, gsdf_lattice *out) {
    if (!bbmin || !bbmax || !out) return fail(GSDF_EINVAL, "gsdf_lattice_from_bounds: NULL argument");
    if (!(res > 0)) return fail(GSDF_EINVAL, "invalid renderer cube resolution");  // flatrenderer.go:38
    for (int a = 0; a < 3; a++) {
        // bb.ScaleCentered(1.01): centre + size*1.01/2 (flatrenderer.go:47-48)
        const float size = bbmax[a] - bbmin[a];
        const float ns = 1.01f * size;
        const float c = bbmin[a] + size * 0.5f;
        const float half = ns * 0.5f;
        const float mn = c - half, mx = c + half;
        const int n = (int)ceilf((mx - mn) / res);  // flatrenderer.go:50-52
        if (n <= 0) return fail(GSDF_ERES, "resolution not fine enough for marching cubes");
        out->n[a] = n;
        out->origin[a] = mn;
    }
    out->res = res;
    return 0;
}

int gsdf_octree_levels(const float bbmin[3], const float bbmax[3], float res) {
    if (!bbmin || !bbmax) return fail(GSDF_EINVAL, "gsdf_octree_levels: NULL argument");
    if (!(res > 0) || std::isinf(res)) return fail(GSDF_EINVAL, "invalid renderer cube resolution");  // octreerenderer.go:223
    float longAxis = 0;
    for (int a = 0; a < 3; a++) {
        const float size = bbmax[a] - bbmin[a];
        const float ns = 1.01f * size, c = bbmin[a] + size * 0.5f, half = ns * 0.5f;
        longAxis = fmaxf(longAxis, (c + half) - (c - half));
    }
    const int levels = (int)ceilf(log2f(longAxis / res)) + 1;  // :229-231
    if (levels <= 1) return fail(GSDF_ERES, "resolution not fine enough for marching cubes");
    return levels;
}

static Lat make_lat(const gsdf_lattice *lat, int k0, int k1, int pitch, bool vec) {
    Lat L;
    L.ox = lat->origin[0]; L.oy = lat->origin[1]; L.oz = lat->origin[2]; L.res = lat->res;
    L.nx = lat->n[0]; L.ny = lat->n[1]; L.nz = lat->n[2];
    L.k0 = k0; L.nk = k1 - k0;
    L.nqx = (lat->n[0] + 1 + 3) / 4;
    L.pitch = pitch;
    L.vec = vec ? 1 : 0;
    return L;
}

int gsdf_grid_eval_device(gsdf_program *p, const gsdf_lattice *lat, int k0, int k1, float *d_dist, void *stream) {
    if (!p || !lat || !d_dist) return fail(GSDF_EINVAL, "gsdf_grid_eval_device: NULL argument");
    if (p->dim != 3) return fail(GSDF_EINVAL, "program is not 3D");
    if (k0 < 0 || k1 > lat->n[2] + 1 || k0 >= k1) return fail(GSDF_EINVAL, "bad corner-plane range [%d,%d)", k0, k1);
    CU(use_device(p->device));
    const int pitch = lat->n[0] + 1;
    const bool vec = (pitch % 4 == 0) && (((uintptr_t)d_dist & 15) == 0);
    GenGrid<4> g{make_lat(lat, k0, k1, pitch, vec), d_dist, nullptr, nullptr};
    const uint64_t nwork = (uint64_t)g.L.nqx * (lat->n[1] + 1) * (k1 - k0);
    return launch_eval<4>(p, g, nwork, stream ? (cudaStream_t)stream : p->stream);
}

int gsdf_grid_eval(gsdf_program *p, const gsdf_lattice *lat, int k0, int k1, float *dist) {
    if (!p || !lat) return fail(GSDF_EINVAL, "gsdf_grid_eval: NULL argument");
    if (k0 < 0 || k1 > lat->n[2] + 1 || k0 >= k1) return fail(GSDF_EINVAL, "bad corner-plane range [%d,%d)", k0, k1);
    CU(use_device(p->device));
    const size_t n = (size_t)(lat->n[0] + 1) * (lat->n[1] + 1) * (k1 - k0);
    int rc = grow(p->d_dist, p->dist_cap, n);
    if (rc) return rc;
    rc = gsdf_grid_eval_device(p, lat, k0, k1, p->d_dist, p->stream);
    if (rc) return rc;
    if (dist) CU(cudaMemcpyAsync(dist, p->d_dist, n * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    p->evals += n;
    return 0;
}

// out: HOST pointer when d_out is NULL (staged through the handle's buffer and copied back), else ignored and the image is
// written straight to the DEVICE buffer d_out on `stream` with no synchronisation.
static int image_run(gsdf_program *p, const float bbmin[2], const float bbmax[2], int w, int h, const gsdf_colorconv *conv, void *out,
                     bool color, void *d_out = nullptr, void *stream = nullptr) {
    if (!p || !bbmin || !bbmax || (!out && !d_out)) return fail(GSDF_EINVAL, "gsdf_image: NULL argument");
    if (p->dim != 2) return fail(GSDF_EINVAL, "program is not 2D");
    if (w <= 0 || h <= 0) return fail(GSDF_EINVAL, "bad image size");
    if (conv && (conv->kind < GSDF_CONV_DEFAULT || conv->kind > GSDF_CONV_HSV_GRADIENT)) return fail(GSDF_EINVAL, "unknown colour conversion %d", conv->kind);
    CU(use_device(p->device));
    const size_t n = (size_t)w * h;
    int rc = d_out ? 0 : grow(p->d_dist, p->dist_cap, n);  // 4 B/pixel either way (float or RGBA8)
    if (rc) return rc;
    float *target = d_out ? static_cast<float *>(d_out) : p->d_dist;
    cudaStream_t st = d_out && stream ? (cudaStream_t)stream : p->stream;
    GenImage g{};
    g.dx = (bbmax[0] - bbmin[0]) / (float)w;  // image.go:85-87
    g.dy = (bbmax[1] - bbmin[1]) / (float)h;
    g.xmin = bbmin[0] + g.dx / 2;
    g.ymax = bbmax[1];                        // un-shifted Max, image.go:92
    g.w = w; g.h = h; g.dist = target;
    g.rgba = color ? reinterpret_cast<uint32_t *>(target) : nullptr;
    g.cc.kind = GSDF_CONV_DEFAULT;
    if (conv) { g.cc.kind = conv->kind; for (int i = 0; i < 7; i++) g.cc.p[i] = conv->p[i]; g.cc.c0 = conv->c0; g.cc.c1 = conv->c1; }
    rc = launch_eval<4>(p, g, (uint64_t)((((w + 3) / 4) + 31) / 32) * ((h + 15) / 16) * 512u, st);
    if (rc) return rc;
    p->evals += n;
    if (d_out) return 0;
    CU(cudaMemcpyAsync(out, p->d_dist, n * 4, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    return 0;
}

int gsdf_image_eval2_device(gsdf_program *p, const float bbmin[2], const float bbmax[2], int w, int h, float *d_dist, void *stream) {
    if (!d_dist || ((uintptr_t)d_dist & 15)) return fail(GSDF_EINVAL, "gsdf_image_eval2_device: d_dist must be a 16-byte aligned device pointer");
    return image_run(p, bbmin, bbmax, w, h, nullptr, nullptr, false, d_dist, stream);
}

int gsdf_image_render2_device(gsdf_program *p, const float bbmin[2], const float bbmax[2], int w, int h, const gsdf_colorconv *conv,
                              uint8_t *d_rgba, void *stream) {
    if (!d_rgba || ((uintptr_t)d_rgba & 15)) return fail(GSDF_EINVAL, "gsdf_image_render2_device: d_rgba must be a 16-byte aligned device pointer");
    return image_run(p, bbmin, bbmax, w, h, conv, nullptr, true, d_rgba, stream);
}

int gsdf_image_eval2(gsdf_program *p, const float bbmin[2], const float bbmax[2], int w, int h, float *dist) {
    return image_run(p, bbmin, bbmax, w, h, nullptr, dist, false);
}

int gsdf_image_render2(gsdf_program *p, const float bbmin[2], const float bbmax[2], int w, int h, const gsdf_colorconv *conv, uint8_t *rgba) {
    return image_run(p, bbmin, bbmax, w, h, conv, rgba, true);
}

int gsdf_colorconv_inigo_quilez(float characteristic_distance, gsdf_colorconv *out) {
    if (!out) return fail(GSDF_EINVAL, "gsdf_colorconv_inigo_quilez: NULL argument");
    *out = gsdf_colorconv{};
    out->kind = GSDF_CONV_INIGO_QUILEZ;
    out->p[0] = 1.f / characteristic_distance;  // color.go:22
    return 0;
}

// gsdfaux/color.go:192-217 on float32
static void rgb_to_hsv(float r, float g, float b, float &h, float &s, float &v) {
    const float xmax = std::max(r, std::max(g, b)), xmin = std::min(r, std::min(g, b));
    const float c = xmax - xmin;
    v = xmax;
    h = 0.f; s = 0.f;
    if (c == 0.f) h = 0.f;
    else if (v == r) h = (g - b) / (c * 6.f);
    else if (v == g) h = (float)(1.0 / 3) + (b - r) / (c * 6.f);
    else if (v == b) h = (float)(2.0 / 3) + (r - g) / (c * 6.f);
    if (h < 0.f) h += 1.f;
    if (xmax > 0.f) s = c / xmax;
}

int gsdf_colorconv_linear_gradient(float gradient_length, uint32_t rgba0, uint32_t rgba1, gsdf_colorconv *out) {
    if (!out) return fail(GSDF_EINVAL, "gsdf_colorconv_linear_gradient: NULL argument");
    *out = gsdf_colorconv{};
    if (rgba0 == 0xff000000u && rgba1 == 0xffffffffu) {  // color.Black -> color.White (color.go:52-54)
        out->kind = GSDF_CONV_BW_LINEAR;
        out->p[0] = gradient_length;
        return 0;
    }
    out->kind = GSDF_CONV_HSV_GRADIENT;
    const uint32_t c[2] = {rgba0, rgba1};
    for (int i = 0; i < 2; i++)  // colorToHSV (color.go:127-130) on the 8-bit channels
        rgb_to_hsv((float)(c[i] & 255u) / 255.f, (float)((c[i] >> 8) & 255u) / 255.f, (float)((c[i] >> 16) & 255u) / 255.f, out->p[3 * i], out->p[3 * i + 1],
                   out->p[3 * i + 2]);
    out->p[6] = gradient_length;
    out->c0 = rgba0;
    out->c1 = rgba1;
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ mesher
struct gsdf_mesher {
    gsdf_program *prog = nullptr;
    gsdf_lattice lat{};
    unsigned flags = 0;
    MeshDims D{};
    float *d_grid = nullptr; size_t grid_cap = 0;
    uint32_t *d_mbits = nullptr; size_t mbits_cap = 0;
    uint32_t *d_list = nullptr; size_t list_cap = 0;
    uint32_t *d_seg = nullptr; size_t seg_cap = 0;
    uint32_t *d_seglist = nullptr; size_t seglist_cap = 0;
    uint8_t *d_segcases = nullptr; size_t segcases_cap = 0;  // 32 case bytes per listed segment (TMA count pass -> emit)
    uint32_t *d_blocksum = nullptr; size_t blocksum_cap = 0;
    unsigned long long *d_scanstate = nullptr; size_t scanstate_cap = 0;
    uint32_t scan_epoch = 0;
    float *d_tris = nullptr; size_t tri_cap = 0;  // in floats
    uint8_t *d_cases = nullptr; size_t cases_cap = 0;
    uint8_t *d_stl = nullptr; size_t stl_cap = 0;
    // device counters: [0] quad list length, [1] overflow flag, [2..3] total triangles (u64), [4] kept blocks
    uint32_t *d_ctr = nullptr;
    uint32_t *h_ctr = nullptr;  // pinned mirror
    CUtensorMap tmap;             // 3-D view of d_grid for the TMA-staged classification
    const float *tmap_grid = nullptr;
    bool use_tma = true;
    uint64_t ntri = 0, evals = 0, pruned = 0, read_pos = 0;
    cudaEvent_t ev[5] = {};
    cudaStream_t copy_stream = nullptr;
    float ms[5] = {};
    // steady-state reruns replay the whole launch sequence (2 memsets + 7 kernels) as ONE CUDA graph
    cudaGraphExec_t gexec = nullptr;
    std::vector<uint8_t> gkey;  // snapshot of every pointer / size the captured launches were built from
    bool allow_graph = true;
    int device = 0;   // device of the program the mesher was created on (destroy must not touch prog: it may be gone)
    uint64_t runs = 0;
    // a render that was enqueued (mesh_run_begin) and not yet finished (mesh_run_end)
    bool pending = false, pend_graph = false, pend_emitted = false;
    MCArgs pendA{};
    unsigned pend_mcgrid = 0;
};

namespace {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_grid_tensor_map(CUtensorMap *out, float *grid, int pitch, int rows, int planes) {
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return fail(GSDF_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
        fn = (encode_tiled_fn)p;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * 4 * (cuuint64_t)rows};  // bytes, dims 1..2
    const cuuint32_t box[3] = {(cuuint32_t)kBoxX, (cuuint32_t)kBoxY, (cuuint32_t)kBoxZ};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, grid, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(GSDF_ECUDA, "cuTensorMapEncodeTiled failed (%d) for a %d x %d x %d lattice", (int)r, pitch, rows, planes);
    return 0;
}

unsigned grid_for(uint64_t items, int per_block, int waves = 8) {
    uint64_t b = (items + per_block - 1) / per_block;
    b = std::min<uint64_t>(b, (uint64_t)g_sms * waves);
    return (unsigned)std::max<uint64_t>(b, 1);
}

int mesh_run_end(gsdf_mesher *m);

// Enqueues one render on the program's stream and returns without waiting (mesh_run_end finishes it).
int mesh_run_begin(gsdf_mesher *m) {
    if (m->pending) { int erc = mesh_run_end(m); if (erc) return erc; }
    gsdf_program *p = m->prog;
    CU(use_device(p->device));
    cudaStream_t st = p->stream;
    if (m->copy_stream) CU(cudaStreamSynchronize(m->copy_stream));  // an earlier async read may still use d_tris
    const MeshDims &D = m->D;
    const bool prune = (m->flags & GSDF_MESH_PRUNE) != 0;
    const int nk = D.cz1 - D.cz0 + 1;
    const uint64_t nquads = (uint64_t)D.nqx * (D.ny + 1) * nk;
    const uint64_t nrows = (uint64_t)D.ny * (D.cz1 - D.cz0);
    const uint64_t ncells = nrows * D.nx;
    if (nquads >= 0xffffffffull) return fail(GSDF_EINVAL, "slab too large: %llu lattice quads (limit 2^32); use more Z-slabs", (unsigned long long)nquads);
    int rc;
    if ((rc = grow(m->d_grid, m->grid_cap, (size_t)D.pitch * (D.ny + 1) * nk))) return rc;
    const uint64_t nseg = nrows * (uint64_t)D.nsx;
    if (nseg >= 0xfff00000ull) return fail(GSDF_EINVAL, "slab too large: %llu cell segments (limit 2^32 - 2^20: grid-stride counters are 32-bit); use more Z-slabs", (unsigned long long)nseg);
    if ((rc = grow(m->d_seg, m->seg_cap, (size_t)nseg))) return rc;
    if ((rc = grow(m->d_seglist, m->seglist_cap, (size_t)nseg))) return rc;
    static const bool pre_classified = getenv("GSDF_EMIT_RECLASSIFY") == nullptr;  // A/B switch: pass 2 classifies again
    if (m->use_tma && pre_classified) {  // one byte per cell at most (every segment listed): never overflows
        if ((rc = grow(m->d_segcases, m->segcases_cap, (size_t)nseg * 32))) return rc;
    }
    const uint64_t nscanblocks = (nseg + kThreads * kScanItems - 1) / (kThreads * kScanItems);
    if ((rc = grow(m->d_blocksum, m->blocksum_cap, (size_t)nscanblocks))) return rc;
    const uint64_t nscantiles = (nseg + kScanTile - 1) / kScanTile;
    if (nscantiles > m->scanstate_cap) {
        if ((rc = grow(m->d_scanstate, m->scanstate_cap, (size_t)nscantiles))) return rc;
        CU(cudaMemsetAsync(m->d_scanstate, 0, m->scanstate_cap * sizeof(unsigned long long), st));
        m->scan_epoch = 0;
    }
    if (++m->scan_epoch >= (1u << 29)) {  // epoch field is 30 bits wide
        CU(cudaMemsetAsync(m->d_scanstate, 0, m->scanstate_cap * sizeof(unsigned long long), st));
        m->scan_epoch = 1;
    }
    if (prune) {
        if ((rc = grow(m->d_mbits, m->mbits_cap, (size_t)D.nwx * D.nby * D.nbz))) return rc;
        if ((rc = grow(m->d_list, m->list_cap, (size_t)nquads))) return rc;
    }
    if (m->flags & GSDF_MESH_KEEP_CASES) {
        if ((rc = grow(m->d_cases, m->cases_cap, (size_t)ncells))) return rc;
    }
    const gsdf_lattice &lat = m->lat;
    MCArgs A;
    A.D = D;
    A.ox = lat.origin[0]; A.oy = lat.origin[1]; A.oz = lat.origin[2]; A.res = lat.res;
    A.cubeDiag = (float)(2 * 1.73205080757) * lat.res;  // flatrenderer.go:202
    A.grid = m->d_grid;
    A.mbits = prune ? m->d_mbits : nullptr;
    CU(cudaGetSymbolAddress((void **)&A.t_ntri, g_mc_ntri));
    CU(cudaGetSymbolAddress((void **)&A.t_tris, g_mc_tris));
    A.segcount = m->d_seg;
    A.tris = m->d_tris;
    A.tri_capacity = m->tri_cap / 9;
    A.cases = (m->flags & GSDF_MESH_KEEP_CASES) ? m->d_cases : nullptr;
    A.overflow = m->d_ctr + 1;
    A.seg_list = m->d_seglist;
    A.seg_count = m->d_ctr + 5;
    A.seg_cases = (m->use_tma && pre_classified) ? m->d_segcases : nullptr;
    const unsigned mcgrid = grid_for(nrows * (uint64_t)((D.nsx + 3) / 4), kThreads / 32, 16);
    if (m->use_tma && m->tmap_grid != m->d_grid) {  // (re)describe the lattice buffer: pitch x (ny+1) x nk floats
        if ((rc = make_grid_tensor_map(&m->tmap, m->d_grid, D.pitch, D.ny + 1, nk))) return rc;
        m->tmap_grid = m->d_grid;
    }
    const bool emitted = m->tri_cap > 0;  // optimistic emit into the existing buffer (steady state: no mid-pipeline host sync)
    static const bool scan3 = getenv("GSDF_SCAN3") != nullptr;  // A/B: the three-kernel scan

    // The launch sequence of one render. stage_events: record the per-stage timing events (eager path only).
    // Programmatic dependent launch between the kernels of the render: every kernel but the first carries the
    // attribute. Stage-timed renders keep plain launches (an event record between two kernels breaks the chain anyway).
    static const bool pdl_on = !(getenv("GSDF_PDL") != nullptr && getenv("GSDF_PDL")[0] == '0');  // default on; GSDF_PDL=0 is the A/B switch
    auto enqueue = [&](bool stage_events, uint32_t epoch) -> int {
    int rc = 0;
    const bool pdl = pdl_on && !stage_events && !scan3 && !(m->flags & GSDF_MESH_KEEP_GRID);
    if (m->flags & GSDF_MESH_KEEP_GRID) CU(cudaMemsetAsync(m->d_grid, 0x7f, (size_t)D.pitch * (D.ny + 1) * nk * sizeof(float), st));
    // counters and look-back scan state are already zero: re-armed by the previous render's k_finish_render (or by the allocation)
    if (prune) {
        GenCenters gc;
        gc.ox = lat.origin[0]; gc.oy = lat.origin[1]; gc.oz = lat.origin[2]; gc.res = lat.res;
        gc.nbx = D.nbx; gc.nby = D.nby; gc.nbz = D.nbz; gc.bz0 = D.bz0;
        const float size = lat.res * 4.0f;          // ms3.Octree.CubeSize of a level-3 cube
        gc.half = size * 0.5f;
        gc.maxDist = size * (float)(1.73205080757 / 2);  // octreerenderer.go:182 with glrender.go:9
        gc.nwx = D.nwx; gc.bits = m->d_mbits; gc.kept = m->d_ctr + 4;
        if ((rc = launch_eval<1>(p, gc, (uint64_t)D.nwx * 32u * D.nby * D.nbz, st))) return rc;
        const uint64_t ncrows = (uint64_t)(D.ny + 1) * nk;
        CU(launch_chain(pdl, k_compact_quads, dim3(grid_for(ncrows, kThreads / 32)), dim3(kThreads), 0, st, D, (const uint32_t *)m->d_mbits, m->d_list, m->d_ctr + 0));
        CU(cudaGetLastError());
    }
    if (stage_events) CU(cudaEventRecord(m->ev[1], st));
    {
        GenGrid<4> g{make_lat(&lat, D.cz0, D.cz0 + nk, D.pitch, true), m->d_grid, prune ? m->d_list : nullptr, prune ? m->d_ctr + 0 : nullptr};
        // with a device-side list length the launch is sized for the worst case; surplus CTAs find no tile and exit
        if ((rc = launch_eval<4>(p, g, nquads, st, pdl && prune))) return rc;
    }
    if (stage_events) CU(cudaEventRecord(m->ev[2], st));
    if (m->use_tma) {
        const uint64_t ntiles = (uint64_t)((D.nsx + 3) / 4) * ((D.ny + kTileY - 1) / kTileY) * (D.cz1 - D.cz0);
        static const bool count_v1 = getenv("GSDF_COUNT_V1") != nullptr;  // A/B switch: one cell per lane, no prefetch
        // test knob: cap the grid so that small, oracle-checked lattices run many tiles per CTA through both stencil buffers
        static const unsigned count_grid_cap = getenv("GSDF_COUNT_GRID") ? (unsigned)std::max(1, atoi(getenv("GSDF_COUNT_GRID"))) : 0u;
        unsigned cgrid = grid_for(ntiles, 1, 16);
        if (count_grid_cap) cgrid = std::min(cgrid, count_grid_cap);
        CU(launch_chain(pdl, count_v1 ? k_mc_count_tma : k_mc_count_tma4, dim3(cgrid), dim3(256), 0, st, m->tmap, A));
    } else {
        CU(launch_chain(pdl, k_mc_count, dim3(mcgrid), dim3(kThreads), 0, st, A));
    }
    CU(cudaGetLastError());
    if (scan3) {
        k_scan_reduce<<<(unsigned)nscanblocks, kThreads, 0, st>>>(m->d_seg, nseg, m->d_blocksum);
        CU(cudaGetLastError());
        k_scan_blocksums<<<1, 1024, 0, st>>>(m->d_blocksum, (uint32_t)nscanblocks, reinterpret_cast<unsigned long long *>(m->d_ctr + 2));
        CU(cudaGetLastError());
        k_scan_apply<<<(unsigned)nscanblocks, kThreads, 0, st>>>(m->d_seg, nseg, m->d_blocksum);
        CU(cudaGetLastError());
    } else {
        CU(launch_chain(pdl, k_scan_lookback, dim3((unsigned)nscantiles), dim3(kThreads), 0, st, m->d_seg, (uint32_t)nseg, m->d_scanstate, m->d_ctr + 6, epoch,
                        reinterpret_cast<unsigned long long *>(m->d_ctr + 2)));
        CU(cudaGetLastError());
    }
    if (stage_events) CU(cudaEventRecord(m->ev[3], st));
    if (emitted) {
        MCArgs E = A;
        E.cases = nullptr;
        CU(launch_chain(pdl, k_mc_emit, dim3(mcgrid), dim3(kThreads), 0, st, E));
        CU(cudaGetLastError());
    }
    {   // publish the counters (cudaMallocHost memory is device-mapped under UVA) and re-arm the state for the next render
        const uint32_t nstate = (uint32_t)nscantiles;
        CU(launch_chain(pdl, k_finish_render, dim3((unsigned)std::min<uint64_t>(std::max<uint64_t>((nstate + 255) / 256, 1), 64)), dim3(256), 0, st,
                        m->d_ctr, (volatile uint32_t *)m->h_ctr, 8, m->d_scanstate, nstate));
        CU(cudaGetLastError());
    }
    return rc;
    };  // enqueue

    // Graph key: everything the captured launches were built from. Any change (buffer regrowth, another program,
    // gsdf_program_update with a different size) re-captures.
    struct GraphKey {
        const void *ptr[10];
        size_t tri_cap;
        ProgView pv;
        unsigned flags;
        int ext, tma;
    } key;
    std::memset(&key, 0, sizeof key);
    const void *kp[10] = {m->d_grid, nullptr, m->d_mbits, m->d_list, m->d_seg, m->d_seglist, m->d_scanstate, m->d_tris, m->d_cases, m->d_segcases};
    std::memcpy(key.ptr, kp, sizeof kp);
    key.tri_cap = m->tri_cap; key.pv = p->pv; key.flags = m->flags; key.ext = p->needs_ext ? 1 : 0; key.tma = m->use_tma ? 1 : 0;
    const bool use_graph = m->allow_graph && !(m->flags & GSDF_MESH_STAGE_TIMING) && emitted && m->runs > 0 && !scan3;
    if (use_graph) {
        if (!m->gexec || m->gkey.size() != sizeof key || std::memcmp(m->gkey.data(), &key, sizeof key) != 0) {
            if (m->gexec) { cudaGraphExecDestroy(m->gexec); m->gexec = nullptr; }
            cudaGraph_t g = nullptr;
            CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            const int erc = enqueue(false, 1u);
            const cudaError_t ce = cudaStreamEndCapture(st, &g);
            if (erc) { if (g) cudaGraphDestroy(g); return erc; }
            if (ce != cudaSuccess) return fail(GSDF_ECUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
            const cudaError_t ie = cudaGraphInstantiate(&m->gexec, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) { m->gexec = nullptr; return fail(GSDF_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); }
            m->gkey.assign(reinterpret_cast<const uint8_t *>(&key), reinterpret_cast<const uint8_t *>(&key) + sizeof key);
        }
        CU(cudaEventRecord(m->ev[0], st));
        CU(cudaGraphLaunch(m->gexec, st));
        m->scan_epoch = 1;  // every render leaves the look-back state zeroed; the graph scans with epoch 1
    } else {
        CU(cudaEventRecord(m->ev[0], st));
        if ((rc = enqueue(true, m->scan_epoch))) return rc;
    }
    CU(cudaEventRecord(m->ev[4], st));
    m->pending = true; m->pend_graph = use_graph; m->pend_emitted = emitted; m->pendA = A; m->pend_mcgrid = mcgrid;
    return 0;
}

// Waits for the enqueued render, reads its counters, re-emits if the triangle buffer was too small, fills the statistics.
int mesh_run_end(gsdf_mesher *m) {
    if (!m->pending) return 0;
    m->pending = false;
    gsdf_program *p = m->prog;
    CU(use_device(p->device));
    cudaStream_t st = p->stream;
    const MeshDims &D = m->D;
    const bool prune = (m->flags & GSDF_MESH_PRUNE) != 0;
    const int nk = D.cz1 - D.cz0 + 1;
    const uint64_t nblocks = (uint64_t)D.nbx * D.nby * D.nbz;
    const bool use_graph = m->pend_graph, emitted = m->pend_emitted;
    MCArgs A = m->pendA;
    const unsigned mcgrid = m->pend_mcgrid;
    int rc;
    CU(cudaEventSynchronize(m->ev[4]));  // the counters were published to m->h_ctr by the last kernel of the sequence
    uint64_t total;
    std::memcpy(&total, m->h_ctr + 2, 8);
    if (!emitted || total * 9 > m->tri_cap) {
        if (m->copy_stream) CU(cudaStreamSynchronize(m->copy_stream));  // a speculative prefix read may be using d_tris
        if ((rc = grow(m->d_tris, m->tri_cap, (size_t)std::max<uint64_t>(total, 1) * 9))) return rc;
        A.tris = m->d_tris;
        A.tri_capacity = m->tri_cap / 9;
        A.cases = nullptr;
        // k_finish_render re-armed the counters already: give the emit its segment-list length back, clear again after
        CU(cudaMemcpyAsync(m->d_ctr + 5, m->h_ctr + 5, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        k_mc_emit<<<mcgrid, kThreads, 0, st>>>(A);
        CU(cudaGetLastError());
        CU(cudaMemsetAsync(m->d_ctr, 0, 8 * sizeof(uint32_t), st));
        CU(cudaEventRecord(m->ev[4], st));
        CU(cudaStreamSynchronize(st));
    }
    m->ntri = total;
    m->read_pos = 0;
    if (prune) {
        m->evals = nblocks + 4ull * m->h_ctr[0];
        m->pruned = (nblocks - m->h_ctr[4]) * 64ull;  // Cube.DecomposesTo(1) of a level-3 cube = 8^2
    } else {
        m->evals = (uint64_t)(D.nx + 1) * (D.ny + 1) * nk;
        m->pruned = 0;
    }
    if (use_graph) { for (int i = 0; i < 4; i++) m->ms[i] = 0.f; }  // stage events are not recorded inside the graph
    else { for (int i = 0; i < 4; i++) cudaEventElapsedTime(&m->ms[i], m->ev[i], m->ev[i + 1]); }
    cudaEventElapsedTime(&m->ms[4], m->ev[0], m->ev[4]);
    m->runs++;
    return 0;
}

int mesh_run(gsdf_mesher *m) {
    int rc = mesh_run_begin(m);
    return rc ? rc : mesh_run_end(m);
}

}  // namespace

extern "C" {

int gsdf_mesh_begin(gsdf_program *p, const gsdf_lattice *lat, int cz0, int cz1, unsigned flags, gsdf_mesher **out) {
    if (!p || !lat || !out) return fail(GSDF_EINVAL, "gsdf_mesh_begin: NULL argument");
    if (p->dim != 3) return fail(GSDF_EINVAL, "program is not 3D");
    if (!(lat->res > 0) || lat->n[0] <= 0 || lat->n[1] <= 0 || lat->n[2] <= 0) return fail(GSDF_ERES, "resolution not fine enough for marching cubes");
    if (cz0 < 0 || cz1 > lat->n[2] || cz0 >= cz1) return fail(GSDF_EINVAL, "bad cell slab [%d,%d)", cz0, cz1);
    int rc = ensure_device();
    if (rc) return rc;
    CU(use_device(p->device));
    gsdf_mesher *m = new gsdf_mesher();
    m->prog = p;
    m->device = p->device;
    m->lat = *lat;
    m->flags = flags;
    m->use_tma = getenv("GSDF_NO_TMA") == nullptr;  // A/B switch for the classification kernel
    m->allow_graph = getenv("GSDF_NO_GRAPH") == nullptr;  // A/B switch: eager launches instead of the CUDA graph
    MeshDims &D = m->D;
    D.nx = lat->n[0]; D.ny = lat->n[1]; D.nz = lat->n[2];
    D.cz0 = cz0; D.cz1 = cz1;
    D.nbx = (D.nx + 3) / 4; D.nby = (D.ny + 3) / 4;
    D.bz0 = cz0 >> 2;
    D.nbz = ((cz1 + 3) >> 2) - D.bz0;
    D.nqx = (D.nx + 1 + 3) / 4;
    D.pitch = D.nqx * 4;
    D.nsx = (D.nx + 31) / 32;
    D.nwx = (D.nbx + 31) / 32;
    cudaError_t e = cudaMalloc((void **)&m->d_ctr, 8 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(m->d_ctr, 0, 8 * sizeof(uint32_t));  // every render leaves them zeroed for the next (k_finish_render)
    if (e == cudaSuccess) e = cudaMallocHost((void **)&m->h_ctr, 8 * sizeof(uint32_t));
    for (int i = 0; i < 5 && e == cudaSuccess; i++) e = cudaEventCreate(&m->ev[i]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { gsdf_mesh_destroy(m); return fail(GSDF_ECUDA, "mesher setup: %s", cudaGetErrorString(e)); }
    rc = mesh_run(m);
    if (rc) { gsdf_mesh_destroy(m); return rc; }
    *out = m;
    return 0;
}

int gsdf_mesh_rerun(gsdf_mesher *m) {
    if (!m) return fail(GSDF_EINVAL, "gsdf_mesh_rerun: NULL mesher");
    return mesh_run(m);
}

int gsdf_mesh_rerun_begin(gsdf_mesher *m) {
    if (!m) return fail(GSDF_EINVAL, "gsdf_mesh_rerun_begin: NULL mesher");
    return mesh_run_begin(m);
}

int gsdf_mesh_rerun_end(gsdf_mesher *m) {
    if (!m) return fail(GSDF_EINVAL, "gsdf_mesh_rerun_end: NULL mesher");
    return mesh_run_end(m);
}

int64_t gsdf_mesh_read_prefix_async(gsdf_mesher *m, float *tri9, size_t ntris) {
    if (!m || (!tri9 && ntris)) return fail(GSDF_EINVAL, "gsdf_mesh_read_prefix_async: NULL argument");
    CU(use_device(m->prog->device));
    const uint64_t n = std::min<uint64_t>(ntris, m->tri_cap / 9);
    if (n == 0) return 0;
    CU(cudaStreamWaitEvent(m->copy_stream, m->ev[4], 0));  // after the emit of the render enqueued last
    CU(cudaMemcpyAsync(tri9, m->d_tris, n * 9 * sizeof(float), cudaMemcpyDeviceToHost, m->copy_stream));
    return (int64_t)n;
}

int gsdf_mesh_set_program(gsdf_mesher *m, gsdf_program *p) {
    if (!m || !p) return fail(GSDF_EINVAL, "gsdf_mesh_set_program: NULL argument");
    if (p->dim != 3) return fail(GSDF_EINVAL, "program is not 3D");
    if (p->device != m->prog->device) return fail(GSDF_EINVAL, "program lives on another device");
    m->prog = p;
    return 0;
}

int64_t gsdf_mesh_read(gsdf_mesher *m, float *tri9, size_t max_tris) {
    if (!m || !tri9) return fail(GSDF_EINVAL, "gsdf_mesh_read: NULL argument");
    if (m->pending) { const int erc = mesh_run_end(m); if (erc) return erc; }
    if (max_tris < 5) return fail(GSDF_ESHORT, "short buffer");  // flatrenderer.go:187
    CU(use_device(m->prog->device));
    const uint64_t left = m->ntri - m->read_pos;
    const uint64_t n = std::min<uint64_t>(left, max_tris);
    if (n == 0) return 0;  // io.EOF
    CU(cudaMemcpy(tri9, m->d_tris + m->read_pos * 9, n * 9 * sizeof(float), cudaMemcpyDeviceToHost));
    m->read_pos += n;
    return (int64_t)n;
}

int64_t gsdf_mesh_read_async(gsdf_mesher *m, float *tri9, size_t max_tris) {
    if (!m || !tri9) return fail(GSDF_EINVAL, "gsdf_mesh_read_async: NULL argument");
    if (m->pending) { const int erc = mesh_run_end(m); if (erc) return erc; }
    if (max_tris < 5) return fail(GSDF_ESHORT, "short buffer");
    CU(use_device(m->prog->device));
    const uint64_t left = m->ntri - m->read_pos;
    const uint64_t n = std::min<uint64_t>(left, max_tris);
    if (n == 0) return 0;
    CU(cudaStreamWaitEvent(m->copy_stream, m->ev[4], 0));  // emit of the last run has finished
    CU(cudaMemcpyAsync(tri9, m->d_tris + m->read_pos * 9, n * 9 * sizeof(float), cudaMemcpyDeviceToHost, m->copy_stream));
    m->read_pos += n;
    return (int64_t)n;
}

int gsdf_mesh_wait(gsdf_mesher *m) {
    if (!m) return fail(GSDF_EINVAL, "gsdf_mesh_wait: NULL mesher");
    CU(use_device(m->prog->device));
    CU(cudaStreamSynchronize(m->copy_stream));
    return 0;
}

int gsdf_mesh_device_triangles(gsdf_mesher *m, const float **d_tri9, uint64_t *ntri) {
    if (!m) return fail(GSDF_EINVAL, "NULL mesher");
    if (m->pending) return fail(GSDF_EINVAL, "a render is in flight on this mesher: call gsdf_mesh_rerun_end first");
    if (d_tri9) *d_tri9 = m->d_tris;
    if (ntri) *ntri = m->ntri;
    return 0;
}

int gsdf_mesh_stats(const gsdf_mesher *m, uint64_t *evals, uint64_t *pruned, uint64_t *tris) {
    if (!m) return fail(GSDF_EINVAL, "NULL mesher");
    if (m->pending) return fail(GSDF_EINVAL, "a render is in flight on this mesher: call gsdf_mesh_rerun_end first");
    if (evals) *evals = m->evals;
    if (pruned) *pruned = m->pruned;
    if (tris) *tris = m->ntri;
    return 0;
}

int gsdf_mesh_cases(gsdf_mesher *m, uint8_t *cases, size_t nbytes) {
    if (!m || !cases) return fail(GSDF_EINVAL, "NULL argument");
    if (m->pending) return fail(GSDF_EINVAL, "a render is in flight on this mesher: call gsdf_mesh_rerun_end first");
    if (!(m->flags & GSDF_MESH_KEEP_CASES)) return fail(GSDF_EINVAL, "mesher was not created with GSDF_MESH_KEEP_CASES");
    const size_t need = (size_t)m->D.nx * m->D.ny * (m->D.cz1 - m->D.cz0);
    if (nbytes != need) return fail(GSDF_ELEN, "cases buffer must be %zu bytes", need);
    CU(use_device(m->prog->device));
    CU(cudaMemcpy(cases, m->d_cases, need, cudaMemcpyDeviceToHost));
    return 0;
}

int gsdf_mesh_grid(gsdf_mesher *m, float *grid, size_t nfloats) {
    if (!m || !grid) return fail(GSDF_EINVAL, "NULL argument");
    if (m->pending) return fail(GSDF_EINVAL, "a render is in flight on this mesher: call gsdf_mesh_rerun_end first");
    if (!(m->flags & GSDF_MESH_KEEP_GRID)) return fail(GSDF_EINVAL, "mesher was not created with GSDF_MESH_KEEP_GRID");
    const MeshDims &D = m->D;
    const size_t rows = (size_t)(D.ny + 1) * (D.cz1 - D.cz0 + 1);
    if (nfloats != rows * (D.nx + 1)) return fail(GSDF_ELEN, "grid buffer must be %zu floats", rows * (D.nx + 1));
    CU(use_device(m->prog->device));
    CU(cudaMemcpy2D(grid, (size_t)(D.nx + 1) * 4, m->d_grid, (size_t)D.pitch * 4, (size_t)(D.nx + 1) * 4, rows, cudaMemcpyDeviceToHost));
    return 0;
}

int gsdf_mesh_timings(const gsdf_mesher *m, float ms[5]) {
    if (!m || !ms) return fail(GSDF_EINVAL, "NULL argument");
    if (m->pending) return fail(GSDF_EINVAL, "a render is in flight on this mesher: call gsdf_mesh_rerun_end first");
    for (int i = 0; i < 5; i++) ms[i] = m->ms[i];
    return 0;
}

void gsdf_mesh_destroy(gsdf_mesher *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaFree(m->d_grid); cudaFree(m->d_mbits); cudaFree(m->d_list); cudaFree(m->d_seg); cudaFree(m->d_seglist); cudaFree(m->d_segcases); cudaFree(m->d_scanstate); cudaFree(m->d_blocksum);
    cudaFree(m->d_tris); cudaFree(m->d_cases); cudaFree(m->d_stl); cudaFree(m->d_ctr);
    if (m->h_ctr) cudaFreeHost(m->h_ctr);
    for (auto &e : m->ev) if (e) cudaEventDestroy(e);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->gexec) cudaGraphExecDestroy(m->gexec);
    delete m;
    (void)cudaGetLastError();  // teardown never leaves a stale (non-sticky) error behind for the next launch check
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ dual contouring
struct gsdf_dualcontour {
    gsdf_program *prog = nullptr;
    int device = 0;
    float bbmin[3], bbmax[3], res = 0;
    int placer = 0, levels = 0;
    int part = 0, nparts = 1;   // this handle owns the part-th of nparts equal ranges of the octree BFS cube order
    uint64_t owned_cubes = 0;
    DCGrid G{};
    float *d_dist = nullptr; size_t dist_cap = 0;
    uint32_t *d_eidx = nullptr; size_t eidx_cap = 0;
    uint32_t *d_cubekey = nullptr; size_t cubekey_cap = 0;
    float4 *d_dc4 = nullptr; size_t dc4_cap = 0;
    float *d_nrm = nullptr; size_t nrm_cap = 0;
    float3 *d_fin = nullptr; size_t fin_cap = 0;
    uint32_t *d_qcount = nullptr; size_t qcount_cap = 0;
    float *d_tris = nullptr; size_t tri_cap = 0;
    unsigned long long *d_scanstate = nullptr; size_t scanstate_cap = 0;
    uint32_t scan_epoch = 0;
    uint32_t *d_ctr = nullptr;            // [0] scan ticket, [2..3] scan total (u64), [4..5] cubes with neighbours (u64)
    uint32_t *h_ctr = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    uint64_t ncubes = 0, ntri = 0, with_nb = 0, evals = 0;
    float ms = 0;
};

namespace {

// Owned key range of part `part` of `nparts` and the box [lo, hi) of cube origins it must evaluate: its run of top-level
// octants grown by one cube on the low side (FinalVertex of the -1 neighbours, dual_contour.go:282-298) and by one cube on
// the high side (the +1 neighbours whose edge data those vertices need), clipped to the grid. Pure host arithmetic (unit-tested without a device).
void dc_part_region(int levels, int part, int nparts, uint32_t keys[2], int32_t box[6]) {
    const int bits = levels - 1, N = 1 << bits;
    const uint64_t ncell = 1ull << (3 * bits);
    keys[0] = (uint32_t)(ncell * (uint64_t)part / (uint64_t)nparts);
    keys[1] = (uint32_t)(ncell * (uint64_t)(part + 1) / (uint64_t)nparts);
    int lo[3] = {N, N, N}, hi[3] = {0, 0, 0};
    if (nparts == 1) { lo[0] = lo[1] = lo[2] = 0; hi[0] = hi[1] = hi[2] = N; }
    else {
        const uint64_t oct = ncell / 8;  // nparts divides 8: the range is a run of top-level octants
        for (uint64_t k = keys[0]; k < keys[1]; k += oct) {
            int i, j, kk;
            dc_unkey((uint32_t)k, bits, i, j, kk);
            const int c[3] = {i, j, kk};
            for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], c[a]); hi[a] = std::max(hi[a], c[a] + N / 2); }
        }
        for (int a = 0; a < 3; a++) { lo[a] = std::max(0, lo[a] - 1); hi[a] = std::min(N, hi[a] + 1); }
    }
    for (int a = 0; a < 3; a++) { box[a] = lo[a]; box[3 + a] = hi[a]; }
}

int dc_scan(gsdf_dualcontour *d, uint32_t *data, uint32_t n, cudaStream_t st) {
    const uint64_t ntiles = ((uint64_t)n + kScanTile - 1) / kScanTile;
    int rc;
    if (ntiles > d->scanstate_cap) {
        if ((rc = grow(d->d_scanstate, d->scanstate_cap, (size_t)ntiles))) return rc;
        CU(cudaMemsetAsync(d->d_scanstate, 0, d->scanstate_cap * sizeof(unsigned long long), st));
        d->scan_epoch = 0;
    }
    if (++d->scan_epoch >= (1u << 29)) {
        CU(cudaMemsetAsync(d->d_scanstate, 0, d->scanstate_cap * sizeof(unsigned long long), st));
        d->scan_epoch = 1;
    }
    CU(cudaMemsetAsync(d->d_ctr, 0, 4 * sizeof(uint32_t), st));
    if (n == 0) return 0;
    k_scan_lookback<<<(unsigned)ntiles, kThreads, 0, st>>>(data, n, d->d_scanstate, d->d_ctr, d->scan_epoch, reinterpret_cast<unsigned long long *>(d->d_ctr + 2));
    CU(cudaGetLastError());
    return 0;
}

int dc_run(gsdf_dualcontour *d) {
    gsdf_program *p = d->prog;
    CU(use_device(p->device));
    cudaStream_t st = p->stream;
    const DCGrid &G = d->G;
    int rc;
    if ((rc = grow(d->d_dist, d->dist_cap, (size_t)G.ncell + 4))) return rc;
    if ((rc = grow(d->d_eidx, d->eidx_cap, (size_t)G.ncell + 8))) return rc;
    CU(cudaEventRecord(d->ev[0], st));
    // Reset: every level-1 cube origin, in octree BFS order (dual_contour.go:37-57)
    uint32_t keys[2];
    int blo[3], bhi[3];
    {
        int32_t box[6];
        dc_part_region(d->levels, d->part, d->nparts, keys, box);
        for (int a = 0; a < 3; a++) { blo[a] = box[a]; bhi[a] = box[3 + a]; }
    }
    const uint32_t key0 = keys[0], key1 = keys[1];
    GenDC g{};
    g.mode = 0; g.G = G; g.dist = d->d_dist;
    for (int a = 0; a < 3; a++) { g.blo[a] = blo[a]; g.bhi[a] = bhi[a]; }
    g.clip = d->nparts > 1 ? 1 : 0;
    if ((rc = launch_eval<4>(p, g, ((uint64_t)G.ncell + 3) / 4, st))) return rc;
    k_dc_flags<<<grid_for(G.ncell, 256), 256, 0, st>>>(d->d_dist, G.ncell, G.res, d->d_eidx);
    CU(cudaGetLastError());
    if ((rc = dc_scan(d, d->d_eidx, G.ncell, st))) return rc;
    CU(cudaMemcpyAsync(d->h_ctr, d->d_ctr, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    unsigned long long tot;
    std::memcpy(&tot, d->h_ctr + 2, 8);
    d->ncubes = tot;
    d->ntri = 0; d->with_nb = 0;
    const uint32_t nc = (uint32_t)d->ncubes;
    const uint64_t norig = (uint64_t)(bhi[0] - blo[0]) * (bhi[1] - blo[1]) * (bhi[2] - blo[2]);
    d->evals = norig + 4ull * nc + (d->placer != GSDF_DC_NAIVE ? 18ull * nc : 0ull);
    if (nc == 0) {
        CU(cudaEventRecord(d->ev[1], st));
        CU(cudaStreamSynchronize(st));
        cudaEventElapsedTime(&d->ms, d->ev[0], d->ev[1]);
        return 0;
    }
    if ((rc = grow(d->d_cubekey, d->cubekey_cap, (size_t)nc))) return rc;
    if ((rc = grow(d->d_dc4, d->dc4_cap, (size_t)nc))) return rc;
    if ((rc = grow(d->d_fin, d->fin_cap, (size_t)nc))) return rc;
    if ((rc = grow(d->d_qcount, d->qcount_cap, (size_t)nc + 8))) return rc;
    k_dc_compact<<<grid_for(G.ncell, 256), 256, 0, st>>>(d->d_dist, d->d_eidx, G.ncell, G.res, d->d_cubekey);
    CU(cudaGetLastError());
    // RenderAll: origin + edge ends (dual_contour.go:85-107)
    g.mode = 1; g.cubekey = d->d_cubekey; g.ncubes = nc; g.dc4 = d->d_dc4;
    if ((rc = launch_eval<4>(p, g, nc, st))) return rc;
    DCArgs A{};
    A.G = G; A.dist = d->d_dist; A.eidx = d->d_eidx; A.cubekey = d->d_cubekey; A.ncubes = nc; A.dc4 = d->d_dc4;
    A.fin = d->d_fin; A.qcount = d->d_qcount; A.placer = d->placer;
    A.with_neighbors = reinterpret_cast<unsigned long long *>(d->d_ctr + 4);
    A.key0 = key0; A.key1 = key1;
    if (d->placer != GSDF_DC_NAIVE) {
        if ((rc = grow(d->d_nrm, d->nrm_cap, (size_t)nc * 9))) return rc;
        const double normStep = d->placer == GSDF_DC_LEAST_SQUARES_CHISELED ? 1e-4 : 2e-8;  // vertexplacement.go:42-46
        float step = (float)normStep;
        step *= 0.5f;  // gleval.go:54
        g.mode = 2; g.step = step; g.nrm = d->d_nrm;
        if ((rc = launch_eval<4>(p, g, (uint64_t)nc * 6, st))) return rc;
        A.nrm = d->d_nrm;
        A.sqrtLambda = d->placer == GSDF_DC_LEAST_SQUARES_CHISELED ? (float)(std::sqrt(1e-5) * normStep) : (float)std::sqrt(1e-5);  // :116-122
    }
    CU(cudaMemsetAsync(d->d_ctr + 4, 0, 2 * sizeof(uint32_t), st));
    k_dc_place<<<(nc + 127) / 128, 128, 0, st>>>(A);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(d->h_ctr + 4, d->d_ctr + 4, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if ((rc = dc_scan(d, d->d_qcount, nc, st))) return rc;
    CU(cudaMemcpyAsync(d->h_ctr, d->d_ctr, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::memcpy(&tot, d->h_ctr + 2, 8);
    const uint64_t nquads = tot;
    std::memcpy(&tot, d->h_ctr + 4, 8);
    d->with_nb = tot;
    d->ntri = 2 * nquads;
    if (nquads) {
        if ((rc = grow(d->d_tris, d->tri_cap, (size_t)nquads * 18))) return rc;
        A.tris = d->d_tris;
        k_dc_emit<<<(nc + 127) / 128, 128, 0, st>>>(A);
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(d->ev[1], st));
    CU(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&d->ms, d->ev[0], d->ev[1]);
    return 0;
}

}  // namespace

extern "C" {

int gsdf_dc_levels(const float bbmin[3], const float bbmax[3], float res, float origin[3]) {
    if (!bbmin || !bbmax) return fail(GSDF_EINVAL, "gsdf_dc_levels: NULL argument");
    if (!(res > 0) || std::isnan(res) || std::isinf(res)) return fail(GSDF_EINVAL, "invalid renderer cube resolution");  // octreerenderer.go:223-225
    const float sub = res / 2;  // dual_contour.go:31-32: bb = Bounds().Add(-res/2) (a translation)
    float mn[3], mx[3];
    for (int a = 0; a < 3; a++) { mn[a] = bbmin[a] + -sub; mx[a] = bbmax[a] + -sub; }
    const float longAxis = std::fmax(mx[0] - mn[0], std::fmax(mx[1] - mn[1], mx[2] - mn[2]));
    const int levels = (int)std::ceil(std::log2(longAxis / res)) + 1;  // octreerenderer.go:229-231
    if (levels <= 1) return fail(GSDF_ERES, "resolution not fine enough for marching cubes");
    if (origin) { origin[0] = mn[0]; origin[1] = mn[1]; origin[2] = mn[2]; }
    return levels;
}

int gsdf_dc_part_region(int levels, int part, int nparts, uint32_t keys[2], int32_t box[6]) {
    if (!keys || !box) return fail(GSDF_EINVAL, "gsdf_dc_part_region: NULL argument");
    if (levels < 2 || levels > 11) return fail(GSDF_EINVAL, "dual contour octree levels must be in [2, 11]");
    if (!(nparts == 1 || nparts == 2 || nparts == 4 || nparts == 8) || part < 0 || part >= nparts)
        return fail(GSDF_EINVAL, "dual contour parts: nparts must be 1, 2, 4 or 8 (runs of top-level octants) and 0 <= part < nparts");
    dc_part_region(levels, part, nparts, keys, box);
    return 0;
}

int gsdf_dc_begin(gsdf_program *p, const float bbmin[3], const float bbmax[3], float res, int placer, gsdf_dualcontour **out) {
    return gsdf_dc_begin_part(p, bbmin, bbmax, res, placer, 0, 1, out);
}

int gsdf_dc_begin_part(gsdf_program *p, const float bbmin[3], const float bbmax[3], float res, int placer, int part, int nparts,
                       gsdf_dualcontour **out) {
    if (!p || !bbmin || !bbmax || !out) return fail(GSDF_EINVAL, "gsdf_dc_begin: NULL argument");
    if (!(nparts == 1 || nparts == 2 || nparts == 4 || nparts == 8) || part < 0 || part >= nparts)
        return fail(GSDF_EINVAL, "dual contour parts: nparts must be 1, 2, 4 or 8 (runs of top-level octants) and 0 <= part < nparts");
    if (p->dim != 3) return fail(GSDF_EINVAL, "program is not 3D");
    if (placer < GSDF_DC_NAIVE || placer > GSDF_DC_LEAST_SQUARES_CHISELED) return fail(GSDF_EINVAL, "nil DualContourer argument to Reset");  // dual_contour.go:28-30
    float org[3];
    const int levels = gsdf_dc_levels(bbmin, bbmax, res, org);
    if (levels < 0) return levels;
    if (levels > 11) return fail(GSDF_EINVAL, "dual contour octree has %d levels (%d^3 cubes); limit is 11 levels", levels, 1 << (levels - 1));
    int rc = ensure_device();
    if (rc) return rc;
    CU(use_device(p->device));
    gsdf_dualcontour *d = new gsdf_dualcontour();
    d->prog = p;
    d->device = p->device;
    for (int a = 0; a < 3; a++) { d->bbmin[a] = bbmin[a]; d->bbmax[a] = bbmax[a]; }
    d->res = res; d->placer = placer; d->levels = levels;
    d->part = part; d->nparts = nparts;
    d->G.ox = org[0]; d->G.oy = org[1]; d->G.oz = org[2]; d->G.res = res;
    d->G.bits = levels - 1;
    d->G.ncell = 1u << (3 * (levels - 1));
    cudaError_t e = cudaMalloc((void **)&d->d_ctr, 8 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&d->h_ctr, 8 * sizeof(uint32_t));
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreate(&d->ev[i]);
    if (e != cudaSuccess) { gsdf_dc_destroy(d); return fail(GSDF_ECUDA, "dual contour setup: %s", cudaGetErrorString(e)); }
    rc = dc_run(d);
    if (rc) { gsdf_dc_destroy(d); return rc; }
    *out = d;
    return 0;
}

int gsdf_dc_rerun(gsdf_dualcontour *d) {
    if (!d) return fail(GSDF_EINVAL, "gsdf_dc_rerun: NULL renderer");
    return dc_run(d);
}

int64_t gsdf_dc_read(gsdf_dualcontour *d, float *tri9, size_t max_tris) {
    if (!d || (!tri9 && max_tris)) return fail(GSDF_EINVAL, "gsdf_dc_read: NULL argument");
    CU(use_device(d->prog->device));
    const uint64_t n = std::min<uint64_t>(d->ntri, max_tris);
    if (n) CU(cudaMemcpy(tri9, d->d_tris, n * 9 * sizeof(float), cudaMemcpyDeviceToHost));
    return (int64_t)n;
}

int gsdf_dc_device_triangles(gsdf_dualcontour *d, const float **d_tri9, uint64_t *ntri) {
    if (!d || !d_tri9 || !ntri) return fail(GSDF_EINVAL, "gsdf_dc_device_triangles: NULL argument");
    *d_tri9 = d->d_tris;
    *ntri = d->ntri;
    return 0;
}

int gsdf_dc_stats(const gsdf_dualcontour *d, uint64_t stats[6]) {
    if (!d || !stats) return fail(GSDF_EINVAL, "gsdf_dc_stats: NULL argument");
    stats[0] = (uint64_t)d->levels; stats[1] = d->ncubes; stats[2] = d->with_nb; stats[3] = d->ntri; stats[4] = d->evals;
    stats[5] = (uint64_t)(d->ms * 1000.f + 0.5f);  /* microseconds of device time */
    return 0;
}

void gsdf_dc_destroy(gsdf_dualcontour *d) {
    if (!d) return;
    cudaSetDevice(d->device);
    cudaFree(d->d_dist); cudaFree(d->d_eidx); cudaFree(d->d_cubekey); cudaFree(d->d_dc4); cudaFree(d->d_nrm); cudaFree(d->d_fin);
    cudaFree(d->d_qcount); cudaFree(d->d_tris); cudaFree(d->d_scanstate); cudaFree(d->d_ctr);
    if (d->h_ctr) cudaFreeHost(d->h_ctr);
    for (auto &e : d->ev) if (e) cudaEventDestroy(e);
    delete d;
    (void)cudaGetLastError();
}

}  // extern "C"

extern "C" {

// ------------------------------------------------------------------------------------------------ STL
static int64_t stl_from_device(const float *d_tri9, uint64_t n, uint8_t *&d_stl, size_t &stl_cap, void *dst, size_t dst_bytes, cudaStream_t st) {
    if (n == 0) return fail(GSDF_EEMPTY, "empty triangle slice");                         // stl.go:16-18
    if (n > 0xffffffffull) return fail(GSDF_EINVAL, "amount of triangles in model exceeds STL design limits");  // stl.go:21-23
    const size_t bytes = 84 + 50 * (size_t)n;
    if (!dst || dst_bytes < bytes) return fail(GSDF_ELEN, "STL buffer needs %zu bytes", bytes);
    // records start 16-byte aligned: 12 bytes of front padding + 84 header bytes = 96
    int rc = grow(d_stl, stl_cap, bytes + 12 + 16);
    if (rc) return rc;
    uint8_t hdr[84] = {0};
    const uint32_t cnt = (uint32_t)n;
    std::memcpy(hdr + 80, &cnt, 4);
    CU(cudaMemcpyAsync(d_stl + 12, hdr, 84, cudaMemcpyHostToDevice, st));
    k_stl_pack<<<grid_for(n, kThreads), kThreads, 0, st>>>(d_tri9, n, d_stl + 96);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(dst, d_stl + 12, bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return (int64_t)bytes;
}

int64_t gsdf_mesh_stl(gsdf_mesher *m, void *dst, size_t dst_bytes) {
    if (!m) return fail(GSDF_EINVAL, "NULL mesher");
    if (m->pending) return fail(GSDF_EINVAL, "a render is in flight on this mesher: call gsdf_mesh_rerun_end first");
    CU(use_device(m->prog->device));
    return stl_from_device(m->d_tris, m->ntri, m->d_stl, m->stl_cap, dst, dst_bytes, m->prog->stream);
}

int64_t gsdf_stl_pack(const float *tri9, size_t n, void *dst, size_t dst_bytes) {
    if (n == 0) return fail(GSDF_EEMPTY, "empty triangle slice");
    if (!tri9) return fail(GSDF_EINVAL, "NULL triangles");
    int rc = ensure_device();
    if (rc) return rc;
    float *d_t = nullptr;
    uint8_t *d_stl = nullptr;
    size_t cap = 0;
    cudaError_t e = cudaMalloc((void **)&d_t, n * 36);
    if (e != cudaSuccess) return fail(GSDF_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
    e = cudaMemcpy(d_t, tri9, n * 36, cudaMemcpyHostToDevice);
    int64_t r = e == cudaSuccess ? stl_from_device(d_t, n, d_stl, cap, dst, dst_bytes, 0) : fail(GSDF_ECUDA, "H2D: %s", cudaGetErrorString(e));
    cudaFree(d_t);
    cudaFree(d_stl);
    return r;
}

}  // extern "C"
