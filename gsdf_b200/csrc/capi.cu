// capi.cu -- implementation of the C ABI declared in include/gsdf_b200.h: errors, devices, programs, Evaluate, the dense
// lattice and the 2-D image entry points. (Interpreter kernels: eval.cu; mesher, multi-device mesher, dual contouring, STL:
// mesher.cu.) Handles own device buffers (grow-never-shrink) and their streams; no CPU fallback exists: every compute
// entry point fails with GSDF_ECUDA when no CUDA device is usable.
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include "internal.cuh"

using namespace gsdfk;
using namespace gsdfi;

namespace gsdfi {

static thread_local std::string g_err;
int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// The default device is per THREAD (gsdf_set_device): two threads that each drive their own GPU never see each other's
// choice. Per-device facts live in a table filled once per device under a lock.
static thread_local int t_device = 0;
int default_device() { return t_device; }

static std::mutex g_dev_mu;
static DevInfo g_dev[64];
static bool g_dev_known[64];

int device_info(int dev, DevInfo *out) {
    if (dev < 0 || dev >= 64) return fail(GSDF_EINVAL, "device %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (!g_dev_known[dev]) {
        int sms = 0, optin = 0;
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CU(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        g_dev[dev].sms = sms;
        g_dev[dev].smem_optin = optin;
        g_dev_known[dev] = true;
    }
    *out = g_dev[dev];
    return 0;
}

int ensure_device(int dev) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return fail(GSDF_ECUDA, "no CUDA device available (%s); libgsdfb200 has no CPU fallback", cudaGetErrorString(e));
    if (dev < 0 || dev >= n) return fail(GSDF_EINVAL, "device %d out of range (have %d)", dev, n);
    CU(use_device(dev));
    return 0;
}

int check_program_blob(const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats, int dim);
void program_add_dependent(gsdf_program *p, cudaEvent_t ev, gsdf_program **ref) {
    std::lock_guard<std::mutex> lk(p->dep_mu);
    p->deps.push_back({ev, ref});
}
void program_remove_dependent(gsdf_program *p, cudaEvent_t ev) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(p->dep_mu);
    for (size_t i = 0; i < p->deps.size(); i++)
        if (p->deps[i].ev == ev) { p->deps.erase(p->deps.begin() + i); break; }
}
int program_quiesce(gsdf_program *p) {
    CU(use_device(p->device));
    CU(cudaStreamSynchronize(p->stream));
    uint64_t dirty = p->user_dirty.exchange(0);
    for (int i = 0; i < kSchedRing; i++)
        if ((dirty >> i) & 1u) CU(cudaEventSynchronize(p->user_ev[i]));
    std::vector<gsdf_program::Dependent> deps;
    {
        std::lock_guard<std::mutex> lk(p->dep_mu);
        deps = p->deps;
    }
    for (const auto &d : deps) CU(cudaEventSynchronize(d.ev));  // an event never recorded is complete
    return 0;
}

}  // namespace gsdfi

namespace {

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
            if (fl) return fail(GSDF_EPROGRAM, "instruction %u: radius-reuse flags, but this library was built with -DGSDF_NO_RXY (set GSDF_RXY=0 in the flattener's environment)", n);
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
        case GSDF_OP_OFFSET: case GSDF_OP_ANNULUS: case GSDF_OP_MULDIST: case GSDF_OP_SHELL_EXIT: case GSDF_OP_BBOX_GUARD2D: case GSDF_OP_MIN_CONST:
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
const char *gsdf_version(void) { return "gsdf-b200 0.2 (sm_100a) +rxy"; }  // with the radius slot (gsdf_program.h)
#else
const char *gsdf_version(void) { return "gsdf-b200 0.2 (sm_100a) -rxy"; }
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
    t_device = device;
    return ensure_device(device);
}

void *gsdf_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0) bytes = 1;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        (void)cudaGetLastError();
        fail(GSDF_ENOMEM, "cudaHostAlloc(%zu bytes) failed", bytes);
        return nullptr;
    }
    return p;
}
void gsdf_host_free(void *p) {
    if (p) cudaFreeHost(p);
    (void)cudaGetLastError();
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

// Operand tables the library derives from a validated program before it goes to the device (the flattened format and its
// producers -- flatten.cpp, the Go flattener -- stay as they are). circarray (cpu_evaluators.go:1056-1078) rotates every point
// by angle*i0 and angle*i1, i0, i1 integers in [0, ncirc]: Sincos of those ncirc+1 angles is computed here once, with the
// same math32 restatement the kernels run (math32.cuh, host side, -ffp-contract=off), appended to the side buffer as
// (sin, cos) pairs, and the table's position (in float4 units, +1) goes into the unused fourth operand word of CIRC_ENTER.
// The kernel then loads two pairs instead of running two Sincos per point; operands outside the table (NaN) take the
// computed path. Word 3 == 0 (every blob that did not pass through here) means no table.
static void augment_program(const gsdf_program_header &h, const uint32_t *chunks, const float *aux, size_t aux_floats, std::vector<uint32_t> &c2,
                            std::vector<float> &a2) {
    c2.assign(chunks, chunks + (size_t)h.nchunks * 4);
    a2.assign(aux, aux + aux_floats);
    for (uint32_t pc = 0; pc < h.nchunks;) {
        const uint32_t op = c2[4 * pc] & 0xff, len = (c2[4 * pc] >> 8) & 0xff;
        if (op == GSDF_OP_CIRC_ENTER && len >= 2 && pc + 1 < h.nchunks) {
            float angle, ncirc;
            std::memcpy(&angle, &c2[4 * (pc + 1)], 4);
            std::memcpy(&ncirc, &c2[4 * (pc + 1) + 1], 4);
            c2[4 * (pc + 1) + 3] = 0u;  // the word belongs to the library: whatever a blob carries there never reaches the kernel
            if (ncirc >= 1.0f && ncirc <= 4096.0f && ncirc == (float)(int)ncirc) {
                const int n = (int)ncirc;
                while (a2.size() & 3) a2.push_back(0.0f);
                c2[4 * (pc + 1) + 3] = (uint32_t)(a2.size() / 4) + 1u;
                for (int i = 0; i <= n; i++) {
                    float sn, cs;
                    m32::sincos(m32::mul(angle, (float)i), sn, cs);
                    a2.push_back(sn); a2.push_back(cs);
                }
                while (a2.size() & 3) a2.push_back(0.0f);
            }
        }
        pc += len ? len : 1;
    }
}

// copies chunks + aux into p->d_blob (growing it if needed) and refreshes the kernel-side view
static int upload_blob(gsdf_program *p, const gsdf_program_header &h, const uint32_t *chunks_in, const float *aux_in, size_t aux_floats_in) {
    std::vector<uint32_t> c2;
    std::vector<float> a2;
    augment_program(h, chunks_in, aux_in, aux_floats_in, c2, a2);
    const uint32_t *chunks = c2.data();
    const float *aux = a2.data();
    const size_t aux_floats = a2.size();
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
    program_note_structure(p, h, chunks);
    return 0;
}

int64_t gsdf_program_device_image(const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats, void *image, size_t image_bytes) {
    gsdf_program_header h;
    const uint32_t *chunks = nullptr;
    const int rc = parse_blob(blob, blob_bytes, aux, aux_floats, h, chunks);
    if (rc) return rc;
    std::vector<uint32_t> c2;
    std::vector<float> a2;
    augment_program(h, chunks, aux, aux_floats, c2, a2);
    const size_t need = c2.size() * 4 + a2.size() * 4;
    if (image) {
        if (image_bytes < need) return fail(GSDF_ESHORT, "device image needs %zu bytes, %zu given", need, image_bytes);
        std::memcpy(image, c2.data(), c2.size() * 4);
        if (!a2.empty()) std::memcpy(static_cast<uint8_t *>(image) + c2.size() * 4, a2.data(), a2.size() * 4);
    }
    return (int64_t)need;
}

int gsdf_program_create_on(int device, const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats, gsdf_program **out) {
    if (!out) return fail(GSDF_EINVAL, "gsdf_program_create: out is NULL");
    gsdf_program_header h;
    const uint32_t *chunks = nullptr;
    int rc = parse_blob(blob, blob_bytes, aux, aux_floats, h, chunks);
    if (rc) return rc;
    rc = ensure_device(device);
    if (rc) return rc;
    DevInfo di;
    if ((rc = device_info(device, &di))) return rc;
    gsdf_program *p = new gsdf_program();
    p->device = device;
    p->sms = di.sms;
    cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_sched, 2 * kSchedRing * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(p->d_sched, 0, 2 * kSchedRing * sizeof(uint32_t));
    for (int i = 0; i < kSchedRing && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&p->user_ev[i], cudaEventDisableTiming);
    if (e != cudaSuccess) { gsdf_program_destroy(p); return fail(GSDF_ECUDA, "program setup: %s", cudaGetErrorString(e)); }
    p->pv.sched = p->d_sched;
    rc = upload_blob(p, h, chunks, aux, aux_floats);
    if (rc) { gsdf_program_destroy(p); return rc; }
    *out = p;
    return 0;
}

int gsdf_program_create(const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats, gsdf_program **out) {
    return gsdf_program_create_on(default_device(), blob, blob_bytes, aux, aux_floats, out);
}

int gsdf_program_update(gsdf_program *p, const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats) {
    if (!p) return fail(GSDF_EINVAL, "gsdf_program_update: NULL program");
    gsdf_program_header h;
    const uint32_t *chunks = nullptr;
    int rc = parse_blob(blob, blob_bytes, aux, aux_floats, h, chunks);
    if (rc) return rc;
    if ((int)h.dim != p->dim) return fail(GSDF_EINVAL, "cannot change a %dD program into a %dD one", p->dim, (int)h.dim);
    if ((rc = program_quiesce(p))) return rc;  // nothing on any stream may still be reading the old program
    return upload_blob(p, h, chunks, aux, aux_floats);
}

void gsdf_program_destroy(gsdf_program *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) (void)program_quiesce(p);
    for (const auto &d : p->deps) if (d.ref) *d.ref = nullptr;  // renderers that outlive their program fail cleanly instead of dangling
    for (auto &lane : p->slot) for (auto &s : lane) {
        if (s.st) cudaStreamDestroy(s.st);
        if (s.done) cudaEventDestroy(s.done);
        if (s.h_pos) cudaFreeHost(s.h_pos);
        if (s.h_dist) cudaFreeHost(s.h_dist);
        cudaFree(s.d_pos);
        cudaFree(s.d_dist);
    }
    for (auto &e : p->user_ev) if (e) cudaEventDestroy(e);
    if (p->upload_ev) cudaEventDestroy(p->upload_ev);
    if (p->h_blob) cudaFreeHost(p->h_blob);
    if (p->stream) cudaStreamDestroy(p->stream);
    cudaFree(p->d_blob);
    cudaFree(p->d_sched);
    cudaFree(p->d_pos);
    cudaFree(p->d_dist);
    delete p;
    (void)cudaGetLastError();
}

uint64_t gsdf_program_evaluations(const gsdf_program *p) { return p ? p->evals.load() : 0; }

// Checks that a caller-supplied device pointer lives on the program's device: a tensor from another GPU would fault the
// context (sticky error) instead of failing the call.
static int check_device_pointer(const gsdf_program *p, const void *ptr, const char *what) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    if (at.type == cudaMemoryTypeDevice && at.device != p->device)
        return fail(GSDF_EINVAL, "%s lives on device %d, the program on device %d", what, at.device, p->device);
    return 0;
}

// a launch on a caller's stream: remember it so that gsdf_program_update / destroy can wait for it
static int note_user_stream(gsdf_program *p, cudaStream_t st) {
    if (st == p->stream) return 0;
    const int slot = (int)(p->sched_next.load(std::memory_order_relaxed) % (uint32_t)kSchedRing);
    CU(cudaEventRecord(p->user_ev[slot], st));
    p->user_dirty.fetch_or(1ull << slot);
    return 0;
}

static int eval_device(gsdf_program *p, const float *d_pos, float *d_dist, size_t n, cudaStream_t st, int dim) {
    const int src = dim == 3 ? launch_stream3(p, d_pos, d_dist, (uint64_t)n, st) : launch_stream2(p, d_pos, d_dist, (uint64_t)n, st);
    if (src <= 0) return src;
    const int vec = (((uintptr_t)d_pos | (uintptr_t)d_dist) & 15) == 0 ? 1 : 0;
    if (dim == 3) return launch_points3(p, GenPoints3{d_pos, d_dist, (uint64_t)n, vec}, (n + 3) / 4, st, nullptr);
    return launch_points2(p, GenPoints2{d_pos, d_dist, (uint64_t)n, vec}, (n + 3) / 4, st, nullptr);
}

int gsdf_eval3_device(gsdf_program *p, const float *d_pos, float *d_dist, size_t n, void *stream) {
    if (!p || !d_pos || !d_dist) return fail(GSDF_EINVAL, "gsdf_eval3_device: NULL argument");
    if (p->dim != 3) return fail(GSDF_EINVAL, "program is not 3D");
    if (n == 0) return fail(GSDF_EEMPTY, "empty buffers");
    int rc;
    if ((rc = check_device_pointer(p, d_pos, "pos")) || (rc = check_device_pointer(p, d_dist, "dist"))) return rc;
    CU(use_device(p->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
    if ((rc = eval_device(p, d_pos, d_dist, n, st, 3))) return rc;
    p->evals += n;
    return note_user_stream(p, st);
}

int gsdf_eval2_device(gsdf_program *p, const float *d_pos, float *d_dist, size_t n, void *stream) {
    if (!p || !d_pos || !d_dist) return fail(GSDF_EINVAL, "gsdf_eval2_device: NULL argument");
    if (p->dim != 2) return fail(GSDF_EINVAL, "program is not 2D");
    if (n == 0) return fail(GSDF_EEMPTY, "empty buffers");
    int rc;
    if ((rc = check_device_pointer(p, d_pos, "pos")) || (rc = check_device_pointer(p, d_dist, "dist"))) return rc;
    CU(use_device(p->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
    if ((rc = eval_device(p, d_pos, d_dist, n, st, 2))) return rc;
    p->evals += n;
    return note_user_stream(p, st);
}

static bool is_pinned(const void *ptr) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// gleval.SDF3.Evaluate on host slices, pipelined. Chunk i travels host -> device, through the kernel and device -> host on
// slot stream i % 3; the three slots overlap one another, so in steady state the PCIe link carries the next chunk in and
// the previous chunk out while the SMs interpret the current one. Pinned caller memory is the DMA source / target itself;
// other memory (a Go slice, a numpy array) goes through the slot's pinned staging, and because one host thread copies at
// ~8 GB/s while the link moves ~55 GB/s, large pageable batches are dealt to up to four lanes -- host threads that live for
// the call, each with its own three slots, each running the same pipeline over every fourth chunk.
constexpr int kEvalLanes = 4;

// chunks c = first, first + stride, ... of the batch through the three slots of one lane
static int eval_host_lane(gsdf_program *p, gsdf_program::EvalSlot *slots, const float *pos, float *dist, size_t n, int dim, size_t chunk,
                          size_t first, size_t stride, bool pin_in, bool pin_out) {
    CU(use_device(p->device));
    const size_t nchunks = (n + chunk - 1) / chunk;
    struct Pending { size_t off = 0, cnt = 0; bool live = false; } pend[3];
    int rc = 0;
    auto drain = [&](int s) -> int {  // wait for the slot's previous chunk and hand its distances to the caller
        gsdf_program::EvalSlot &S = slots[s];
        if (!pend[s].live) return 0;
        CU(cudaEventSynchronize(S.done));
        if (!pin_out) std::memcpy(dist + pend[s].off, S.h_dist, pend[s].cnt * sizeof(float));
        pend[s].live = false;
        return 0;
    };
    size_t k = 0;
    for (size_t c = first; c < nchunks && !rc; c += stride, k++) {
        const int s = (int)(k % 3);
        gsdf_program::EvalSlot &S = slots[s];
        if ((rc = drain(s))) break;
        const size_t off = c * chunk, cnt = std::min(chunk, n - off);
        if (!S.st) {
            CU(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming));
        }
        if (S.cap < cnt) {
            if (S.h_pos) cudaFreeHost(S.h_pos);
            if (S.h_dist) cudaFreeHost(S.h_dist);
            cudaFree(S.d_pos); cudaFree(S.d_dist);
            S.h_pos = S.h_dist = S.d_pos = S.d_dist = nullptr; S.cap = 0;
            cudaError_t e = cudaMalloc((void **)&S.d_pos, cnt * 3 * sizeof(float));
            if (e == cudaSuccess) e = cudaMalloc((void **)&S.d_dist, cnt * sizeof(float));
            if (e == cudaSuccess) e = cudaHostAlloc((void **)&S.h_pos, cnt * 3 * sizeof(float), cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaHostAlloc((void **)&S.h_dist, cnt * sizeof(float), cudaHostAllocDefault);
            if (e != cudaSuccess) return fail(GSDF_ENOMEM, "Evaluate staging (%zu points): %s", cnt, cudaGetErrorString(e));
            S.cap = cnt;
        }
        const float *src = pos + off * dim;
        if (!pin_in) { std::memcpy(S.h_pos, src, cnt * dim * sizeof(float)); src = S.h_pos; }
        CU(cudaMemcpyAsync(S.d_pos, src, cnt * dim * sizeof(float), cudaMemcpyHostToDevice, S.st));
        if ((rc = eval_device(p, S.d_pos, S.d_dist, cnt, S.st, dim))) break;
        CU(cudaMemcpyAsync(pin_out ? dist + off : S.h_dist, S.d_dist, cnt * sizeof(float), cudaMemcpyDeviceToHost, S.st));
        CU(cudaEventRecord(S.done, S.st));
        pend[s].off = off; pend[s].cnt = cnt; pend[s].live = true;
    }
    for (size_t j = 0; j < 3; j++) {  // oldest first
        const int s = (int)((k + j) % 3);
        const int drc = drain(s);
        if (!rc) rc = drc;
    }
    if (rc)
        for (int s = 0; s < 3; s++) if (slots[s].st) cudaStreamSynchronize(slots[s].st);
    return rc;
}

static int eval_host(gsdf_program *p, const float *pos, float *dist, size_t n, int dim) {
    if (!p || !pos || !dist) return fail(GSDF_EINVAL, "gsdf_eval: NULL argument");
    if (p->dim != dim) return fail(GSDF_EINVAL, "program is %dD, called as %dD", p->dim, dim);
    if (n == 0) return fail(GSDF_EEMPTY, "empty buffers");
    CU(use_device(p->device));
    const char *ce = getenv("GSDF_EVAL_CHUNK");  // test / tuning knobs (read per call: a call costs microseconds at least)
    const size_t chunk_env = ce ? (size_t)atoll(ce) : 0;
    const char *le = getenv("GSDF_EVAL_LANES");
    // chunks of 256 Ki points (3 MB in, 1 MB out): long enough for the link to stream, short enough that ramp-up and
    // drain of the pipeline stay small; a batch smaller than two chunks goes as one
    size_t chunk = chunk_env ? chunk_env : (size_t)256 << 10;
    chunk = std::max<size_t>((chunk + 2047) & ~(size_t)2047, 2048);  // whole tiles of the streaming kernel
    if (n < 2 * chunk) chunk = n;
    const bool pin_in = is_pinned(pos), pin_out = is_pinned(dist);
    const size_t nchunks = (n + chunk - 1) / chunk;
    int lanes = 1;
    if (!(pin_in && pin_out) && nchunks >= 8) lanes = (int)std::min<size_t>(kEvalLanes, nchunks / 2);  // host copies are the bottleneck
    if (le) lanes = std::max(1, std::min(kEvalLanes, atoi(le)));
    lanes = (int)std::min<size_t>((size_t)lanes, nchunks);
    int rc = 0;
    if (lanes <= 1) {
        rc = eval_host_lane(p, p->slot[0], pos, dist, n, dim, chunk, 0, 1, pin_in, pin_out);
    } else {
        int lrc[kEvalLanes] = {0, 0, 0, 0};
        std::string lerr[kEvalLanes];
        std::vector<std::thread> th;
        for (int l = 1; l < lanes; l++)
            th.emplace_back([&, l] {
                lrc[l] = eval_host_lane(p, p->slot[l], pos, dist, n, dim, chunk, (size_t)l, (size_t)lanes, pin_in, pin_out);
                if (lrc[l]) lerr[l] = g_err;  // the message lives in this thread's error slot
            });
        lrc[0] = eval_host_lane(p, p->slot[0], pos, dist, n, dim, chunk, 0, (size_t)lanes, pin_in, pin_out);
        for (auto &t : th) t.join();
        rc = lrc[0];
        for (int l = 1; l < lanes && !rc; l++)
            if (lrc[l]) rc = fail(lrc[l], "%s", lerr[l].c_str());
    }
    if (rc) return rc;
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

int gsdf_grid_eval_device(gsdf_program *p, const gsdf_lattice *lat, int k0, int k1, float *d_dist, void *stream) {
    if (!p || !lat || !d_dist) return fail(GSDF_EINVAL, "gsdf_grid_eval_device: NULL argument");
    if (p->dim != 3) return fail(GSDF_EINVAL, "program is not 3D");
    if (k0 < 0 || k1 > lat->n[2] + 1 || k0 >= k1) return fail(GSDF_EINVAL, "bad corner-plane range [%d,%d)", k0, k1);
    int rc;
    if ((rc = check_device_pointer(p, d_dist, "dist"))) return rc;
    CU(use_device(p->device));
    const int pitch = lat->n[0] + 1;
    const bool vec = (pitch % 4 == 0) && (((uintptr_t)d_dist & 15) == 0);
    GenGrid<4> g{make_lat(lat, k0, k1, pitch, vec), d_dist, nullptr, nullptr};
    const uint64_t nwork = (uint64_t)g.L.nqx * (lat->n[1] + 1) * (k1 - k0);
    cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
    if ((rc = launch_grid4(p, g, nwork, st, false, nullptr))) return rc;
    p->evals += (uint64_t)(lat->n[0] + 1) * (lat->n[1] + 1) * (k1 - k0);
    return note_user_stream(p, st);
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
    int rc = d_out ? check_device_pointer(p, d_out, "image") : grow(p->d_dist, p->dist_cap, n);  // 4 B/pixel either way (float or RGBA8)
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
    rc = launch_image(p, g, GenImage::items_for(w, h), st, nullptr);
    if (rc) return rc;
    p->evals += n;
    if (d_out) return note_user_stream(p, st);
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

namespace gsdfi {
int program_update_async(gsdf_program *p, const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats) {
    gsdf_program_header h;
    const uint32_t *chunks_in = nullptr;
    int rc = parse_blob(blob, blob_bytes, aux, aux_floats, h, chunks_in);
    if (rc) return rc;
    if ((int)h.dim != p->dim) return fail(GSDF_EINVAL, "cannot change a %dD program into a %dD one", p->dim, (int)h.dim);
    std::vector<uint32_t> c2;
    std::vector<float> a2;
    augment_program(h, chunks_in, aux, aux_floats, c2, a2);  // (same tables as upload_blob: the layout comparison below is on the augmented sizes)
    const uint32_t *chunks = c2.data();
    const float *aux2 = a2.data();
    const size_t prog_bytes = (size_t)h.nchunks * 16, aux_bytes = a2.size() * 4;
    bool ext = false;
    for (uint32_t pc = 0; pc < h.nchunks;) {
        const uint32_t op = chunks[4 * pc] & 0xff, len = (chunks[4 * pc] >> 8) & 0xff;
        if (op == GSDF_OP_ELLIPSE2D || op == GSDF_OP_BEZIERQ2D) ext = true;
        pc += len ? len : 1;
    }
    const bool same_layout = p->d_blob && prog_bytes == p->pv.prog_bytes && aux_bytes == p->pv.aux_bytes && h.dstack == p->pv.dslots &&
                             h.pstack == p->pv.pslots && ext == p->needs_ext;
    if (!same_layout) return gsdf_program_update(p, blob, blob_bytes, aux, aux_floats);
    CU(use_device(p->device));
    if (p->h_blob_cap < prog_bytes + aux_bytes) {
        if (p->h_blob) cudaFreeHost(p->h_blob);
        p->h_blob = nullptr; p->h_blob_cap = 0;
        CU(cudaHostAlloc((void **)&p->h_blob, 2 * (prog_bytes + aux_bytes) + 64, cudaHostAllocDefault));
        p->h_blob_cap = 2 * (prog_bytes + aux_bytes) + 64;
    }
    if (!p->upload_ev) CU(cudaEventCreateWithFlags(&p->upload_ev, cudaEventDisableTiming));
    if (p->upload_ev_recorded) CU(cudaEventSynchronize(p->upload_ev));  // the staging buffer is free again (normally long done)
    std::memcpy(p->h_blob, chunks, prog_bytes);
    if (aux_bytes) std::memcpy(p->h_blob + prog_bytes, aux2, aux_bytes);
    {   // device-side: nothing that still reads the old program may be overtaken
        std::lock_guard<std::mutex> lk(p->dep_mu);
        for (const auto &d : p->deps) CU(cudaStreamWaitEvent(p->stream, d.ev, 0));
    }
    CU(cudaMemcpyAsync(p->d_blob, p->h_blob, prog_bytes + aux_bytes, cudaMemcpyHostToDevice, p->stream));
    CU(cudaEventRecord(p->upload_ev, p->stream));
    p->upload_ev_recorded = true;
    p->ninstr = h.ninstr;
    program_note_structure(p, h, chunks);
    return 0;
}

void program_note_structure(gsdf_program *p, const gsdf_program_header &h, const uint32_t *chunks) {
    p->h_words.assign(chunks, chunks + (size_t)h.nchunks * 4);
    p->skey = program_structure_key(h, chunks, p->needs_ext, p->pv.stage_aux != 0);
    if (p->jit && p->jit->key != p->skey) p->jit.reset();  // (the interpreter runs until gsdf_program_specialize is called again)
}

// full validation of a flattened program without touching a device (gsdf_multi_update keeps the blob for its workers)
int check_program_blob(const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats, int dim) {
    gsdf_program_header h;
    const uint32_t *chunks = nullptr;
    const int rc = parse_blob(blob, blob_bytes, aux, aux_floats, h, chunks);
    if (rc) return rc;
    if ((int)h.dim != dim) return fail(GSDF_EINVAL, "program is %dD, expected %dD", (int)h.dim, dim);
    return 0;
}
}  // namespace gsdfi
