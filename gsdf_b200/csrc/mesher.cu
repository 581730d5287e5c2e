// mesher.cu -- the mesh stage of the C ABI: gsdf_mesher (one Z-slab of a lattice on one device), gsdf_multimesher (one
// lattice over several slabs and devices from one process), dual contouring and STL packing.
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include "internal.cuh"
#include "mc_kernels.cuh"
#include "mc_tile5.cuh"
#include "mc_block.cuh"
#include "dualcontour.cuh"

using namespace gsdfk;
using namespace gsdfi;

constexpr int kMeshStamps = 8;
constexpr int kMeshCtr = 32;  // words of a mesher's counter block (see gsdf_mesher::d_ctr)
#define GSDF_MESH_PLAN_GIVEN 0u

// ------------------------------------------------------------------------------------------------ mesher
struct gsdf_mesher {
    gsdf_program *prog = nullptr;
    gsdf_lattice lat{};
    unsigned flags = 0;
    MeshDims D{};
    float *d_grid = nullptr; size_t grid_cap = 0;
    uint32_t *d_mbits = nullptr;  // = d_lbits[nlevels-1]: the level-3 bit rows the marching-cubes stage reads
    // coarse-to-fine prune plan (gsdf_prune_plan): per level its geometry on this slab and its bit rows
    gsdf_prune_plan plan{};
    PruneLevel lev[GSDF_PRUNE_MAX_LEVELS]{};
    uint32_t *d_lbits[GSDF_PRUNE_MAX_LEVELS] = {}; size_t lbits_cap[GSDF_PRUNE_MAX_LEVELS] = {};
    // plan ending with level 2 (PruneFine, generators.cuh): nl3 = levels down to level 3, then the 2-cell cubes
    bool fine = false; int nl3 = 0; float half2 = 0.f, maxDist2 = 0.f;
    uint32_t *d_bits2 = nullptr; size_t bits2_cap = 0;        // level-2 rows (E, O words)
    uint8_t *d_childmask = nullptr; size_t childmask_cap = 0; // per level-3 block: surviving children
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
    // device counters: [0] quad list length, [1] overflow flag, [2..3] total triangles (u64), [4] kept level-3 cubes,
    // [5] listed segments, [6] scan ticket, [7] prune-cube centres evaluated, [8..17] work-tile schedulers of this mesher's
    // interpreter launches (one pair per prune level + one for the lattice evaluation: never shared with another launch),
    // [18] scan tiles finished (scan inside the emit pass), [19] kept 2-cell cubes, [31] CTAs of the last kernel that are done
    uint32_t *d_ctr = nullptr;
    uint32_t *h_ctr = nullptr;  // pinned mirror
    // stage stamps (%globaltimer, ns): [0] prune centres, [1] quad compaction, [2] lattice evaluation, [3] classification,
    // [4] scan, [5] emit, [7] finish -- written by the kernels themselves, so they exist inside CUDA-graph replays too
    unsigned long long *d_stamp = nullptr;
    unsigned long long *h_stamp = nullptr;
    CUtensorMap tmap;             // 3-D view of d_grid for the TMA-staged classification (2-plane boxes, one-layer tiles)
    CUtensorMap tmap5;            // the same view with 5-plane boxes (mc_tile5.cuh: four layers per tile)
    // marching-cubes kernels: 2 = kept-block kernels (mc_block.cuh, default), 1 = 4-layer tiles (mc_tile5.cuh),
    // 0 = one-layer tiles + segment work list (mc_kernels.cuh). GSDF_MC=block|tile5|v1 selects (A/B).
    int mc_mode = 2;
    uint32_t *d_blklist = nullptr; size_t blklist_cap = 0;   // kept blocks of the slab
    uint2 *d_blkcnt = nullptr; size_t blkcnt_cap = 0;        // per 32-cell segment: 8 block-slot triangle counts
    CUtensorMap tmapB;            // the lattice with 8x5x5 boxes: one kept block's corner stencil
    uint32_t blk_hint = 0;        // kept blocks of the previous render, rounded up
    bool tile5 = false;           // GSDF_MC_TILE5: the 4-layer-tile kernel pair instead of one-layer tiles + segment work list
    const float *tmap_grid = nullptr;
    bool use_tma = true;
    uint64_t ntri = 0, evals = 0, pruned = 0, read_pos = 0;
    uint32_t quad_hint = 0;       // listed quads of the previous render, rounded up to 4096
    cudaEvent_t ev[5] = {};
    cudaStream_t stream = nullptr;       // the render's stream: every mesher has its own, so slabs of one lattice overlap
    cudaStream_t copy_stream = nullptr;
    float ms[5] = {};
    // steady-state reruns replay the whole launch sequence (2 memsets + 7 kernels) as ONE CUDA graph
    cudaGraphExec_t gexec = nullptr;
    std::vector<uint8_t> gkey;  // snapshot of every pointer / size the captured launches were built from
    bool allow_graph = true;
    bool pdl_chain = true;        // programmatic dependent launch between the kernels of a render (see multi_slab_pdl)
    int device = 0;   // device of the program the mesher was created on (destroy must not touch prog: it may be gone)
    uint64_t runs = 0;
    // a render that was enqueued (mesh_run_begin) and not yet finished (mesh_run_end)
    bool pending = false, pend_graph = false, pend_emitted = false;
    bool last_reemit = false;  // the render just finished had to grow the triangle buffer and emit again (mesh_run_end)
    MCArgs pendA{};
    BlkArgs pendB{};
    unsigned pend_mcgrid = 0, pend_blkgrid = 0;
    bool pend_half = false;  // the pending render's work list holds half-quads (Lat::hq)
};

namespace {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_grid_tensor_map(CUtensorMap *out, float *grid, int pitch, int rows, int planes, int box_z, int box_x = kBoxX, int box_y = kBoxY) {
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
    const cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, (cuuint32_t)box_z};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, grid, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(GSDF_ECUDA, "cuTensorMapEncodeTiled failed (%d) for a %d x %d x %d lattice", (int)r, pitch, rows, planes);
    return 0;
}

int mesh_run_end(gsdf_mesher *m);

// Enqueues one render on the program's stream and returns without waiting (mesh_run_end finishes it).
int mesh_run_begin(gsdf_mesher *m) {
    if (m->pending) { int erc = mesh_run_end(m); if (erc) return erc; }
    gsdf_program *p = m->prog;
    if (!p) return fail(GSDF_EINVAL, "the renderer's program was destroyed");
    CU(use_device(p->device));
    cudaStream_t st = m->stream;
    if (m->copy_stream) CU(cudaStreamSynchronize(m->copy_stream));  // an earlier async read may still use d_tris
    if (p->upload_ev_recorded) CU(cudaStreamWaitEvent(st, p->upload_ev, 0));  // an asynchronous program upload (program_update_async) lands first
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
        for (int li = 0; li < m->nl3; li++) {
            PruneLevel &Lv = m->lev[li];
            if ((rc = grow(m->d_lbits[li], m->lbits_cap[li], (size_t)Lv.nwx * Lv.ncy * Lv.ncz))) return rc;
            Lv.bits = m->d_lbits[li];
        }
        m->d_mbits = m->d_lbits[m->nl3 - 1];
        if (m->fine) {
            if ((rc = grow(m->d_bits2, m->bits2_cap, (size_t)8 * D.nwx * D.nby * D.nbz))) return rc;
            if ((rc = grow(m->d_childmask, m->childmask_cap, (size_t)D.nbx * D.nby * D.nbz))) return rc;
        }
        // (half-quad lists of the two-corners-per-thread lattice kernel: up to two entries per quad)
        if ((rc = grow(m->d_list, m->list_cap, (size_t)nquads * (has_grid2(p) ? 2u : 1u)))) return rc;
    }
    if (m->flags & GSDF_MESH_KEEP_CASES) {
        if ((rc = grow(m->d_cases, m->cases_cap, (size_t)ncells))) return rc;
    }
    // the kept-block kernels pay on pruned lattices; a dense sweep (FlatRenderer: every block kept) is what the row-tile kernels
    // are for (measured: flange@400 dense 0.269 ms with tiles, 0.303 ms with blocks)
    const bool blockmc = m->use_tma && m->mc_mode == 2 && prune;
    const uint64_t nblocks_slab = (uint64_t)D.nbx * D.nby * D.nbz;
    if (blockmc) {
        if ((rc = grow(m->d_blklist, m->blklist_cap, (size_t)nblocks_slab))) return rc;
        if ((rc = grow(m->d_blkcnt, m->blkcnt_cap, (size_t)nseg))) return rc;
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
    A.stamp = nullptr;
    A.fin_ctr = nullptr; A.fin_hctr = nullptr; A.fin_nctr = 0; A.fin_scanstate = nullptr; A.fin_nstate = 0;
    A.fin_dstamp = nullptr; A.fin_hstamp = nullptr; A.fin_nstamp = 0; A.fin_done = nullptr;
    BlkArgs BA{};
    BA.D = D; BA.ox = A.ox; BA.oy = A.oy; BA.oz = A.oz; BA.res = A.res; BA.cubeDiag = A.cubeDiag;
    BA.childmask = (prune && m->fine) ? m->d_childmask : nullptr;
    // (tile sums accumulated by the count pass -- 158 k reductions on 104 addresses -- cost the count pass 13 us and save the
    // scan 2: off; GSDF_TILESUM=1 is the A/B switch)
    static const bool tilesum_on = getenv("GSDF_TILESUM") != nullptr && getenv("GSDF_TILESUM")[0] == '1';
    BA.tilesum = (tilesum_on && nscantiles <= kTileSumMax) ? m->d_scanstate : nullptr;
    BA.mbits = A.mbits; BA.blklist = m->d_blklist; BA.nblk = m->d_ctr + 5; BA.blkcnt = m->d_blkcnt; BA.segoff = m->d_seg;
    BA.t_ntri = A.t_ntri; BA.t_tris = A.t_tris; BA.tris = m->d_tris; BA.tri_capacity = m->tri_cap / 9; BA.cases = A.cases;
    BA.overflow = A.overflow;
    // persistent grid of the block kernels: one warp per kept block, sized from the previous render's count when there is one
    // (launch shapes follow the previous render's counts plus a margin, not the worst case: resident CTAs that find no work
    // still hold registers and shared memory, and under programmatic launch they hold them early -- on a device shared by
    // several slabs that kept the other slabs' kernels out; a render that outgrows the hint is still correct, only slower)
    static const unsigned hint_slack = getenv("GSDF_HINT_SLACK") ? (unsigned)std::max(0, atoi(getenv("GSDF_HINT_SLACK"))) : 8u;  // margin = hint / slack (0: the old 2x)
    auto with_margin = [&](uint64_t hint, uint64_t floor_) { return hint_slack ? hint + hint / hint_slack + floor_ : 2 * hint + floor_; };
    const uint64_t blk_bound = m->runs > 0 ? std::min<uint64_t>(nblocks_slab, with_margin(m->blk_hint, 256)) : nblocks_slab;
    // test knob: cap the grid so that on small, oracle-checked lattices every warp walks many blocks through both stencil buffers
    static const unsigned blk_grid_cap = getenv("GSDF_BLK_GRID") ? (unsigned)std::max(1, atoi(getenv("GSDF_BLK_GRID"))) : 0u;
    static const int blk_waves = getenv("GSDF_BLK_WAVES") ? std::max(1, atoi(getenv("GSDF_BLK_WAVES"))) : 8;  // A/B knob: CTAs per SM of the block kernels' grid
    unsigned blkgrid = grid_for(p->sms, blk_bound, kBlkWarps, blk_waves);
    if (blk_grid_cap) blkgrid = std::min(blkgrid, blk_grid_cap);
    const unsigned mcgrid = grid_for(p->sms, nrows * (uint64_t)((D.nsx + 3) / 4), kThreads / 32, 16);
    if (m->use_tma && m->tmap_grid != m->d_grid) {  // (re)describe the lattice buffer: pitch x (ny+1) x nk floats
        if ((rc = make_grid_tensor_map(&m->tmap, m->d_grid, D.pitch, D.ny + 1, nk, kBoxZ))) return rc;
        if ((rc = make_grid_tensor_map(&m->tmap5, m->d_grid, D.pitch, D.ny + 1, nk, kBox5Z))) return rc;
        if ((rc = make_grid_tensor_map(&m->tmapB, m->d_grid, D.pitch, D.ny + 1, nk, kBlkBoxZ, kBlkBoxX, kBlkBoxY))) return rc;
        m->tmap_grid = m->d_grid;
    }
    const bool emitted = m->tri_cap > 0;  // optimistic emit into the existing buffer (steady state: no mid-pipeline host sync)
    // Points per thread of the lattice evaluation. Four (the throughput form) unless the PREVIOUS render of this handle listed
    // so few quads that one-corner tiles still fit in about one resident wave: then the stage is bound by the latency of a
    // tile (~30 us at four points per thread, a quarter of that at one), which is what thin Z-slabs -- strong scaling, the
    // first slab of a pipelined read-back -- pay for. The persistent grid is sized from the upper bound either way, so a
    // render that lists more than the hint predicted is still correct, only slower.
    int eval_p = 4;
    {
        static const int force_p = getenv("GSDF_EVAL_P") ? atoi(getenv("GSDF_EVAL_P")) : 0;  // A/B switch: 1, 2 or 4
        int slots = 0, cta = kEvalThreads;
        if ((rc = eval_cta_slots(p, &slots, &cta))) return rc;
        if (prune && m->runs > 0 && (uint64_t)m->quad_hint * 4 <= (uint64_t)slots * (uint64_t)cta * 3) eval_p = 1;  // <= 3 short rounds
        if (force_p == 1 && prune && m->runs > 0) eval_p = 1;
        if (force_p == 4) eval_p = 4;
        // the run-time compiled kernels also come with two corners per thread: half the straight-line code per tile and a
        // 40-register cap (12 resident CTAs of 128 threads instead of 9). Measured against four: flange@400 39 vs 39 us, bolt@400 43
        // vs 45, knurled@500 349 vs 368 -- the throughput form whenever the kernels are specialised
        if (eval_p == 4 && force_p != 4 && has_grid2(p)) eval_p = 2;
        if (force_p == 2 && has_grid2(p)) eval_p = 2;
    }
    // Two corners per thread can read a list of half-quads (k_mesh_lists, Lat::hq): boundary quads list their first half only,
    // 5-9 % fewer lattice evaluations. Building the finer list costs the list kernel 2-4 us, so it is used once the previous
    // render listed at least 2^19 quads: flange@400 (0.4 M quads) 40 -> 38 us of evaluation against 16 -> 18 us of prune,
    // knurled@500 (1.25 M) 310 -> 281 against 38 -> 42. GSDF_HALF_QUADS=0 / 1 switch it off / on whatever the size (A/B, tests).
    static const int half_env = getenv("GSDF_HALF_QUADS") ? (getenv("GSDF_HALF_QUADS")[0] == '0' ? 0 : 1) : -1;
    constexpr uint32_t kHalfQuadMin = 1u << 19;
    // (in half mode the hint is half-quads / 2, a little below the quads listed: a render stays in the mode down to 3/4 of
    // the threshold, so that a lattice near it does not flip -- and re-capture its graph -- every other render)
    const bool half_wanted = half_env >= 0 ? half_env == 1 : (m->runs > 0 && m->quad_hint >= (m->pend_half ? kHalfQuadMin / 4 * 3 : kHalfQuadMin));
    const bool half = half_wanted && prune && blockmc && eval_p == 2 && nquads < (1ull << 31) && m->list_cap >= 2 * nquads;
    static const bool scan3 = getenv("GSDF_SCAN3") != nullptr;  // A/B: the three-kernel scan
    // default: the segment scan runs inside the emit pass of the block kernels; GSDF_SCAN_FUSED=0 launches k_scan_seg (A/B)
    static const bool scan_fused_on = !scan3 && !(getenv("GSDF_SCAN_FUSED") != nullptr && getenv("GSDF_SCAN_FUSED")[0] == '0');
    static const uint64_t scan_fused_max = getenv("GSDF_SCAN_FUSED_MAX") ? (uint64_t)atoll(getenv("GSDF_SCAN_FUSED_MAX")) : (1ull << 40);
    const bool scan_fused = scan_fused_on && nscantiles <= scan_fused_max;

    // The launch sequence of one render. stage_events: record the per-stage timing events (eager path only).
    // Programmatic dependent launch between the kernels of the render: every kernel but the first carries the
    // attribute. Stage-timed renders keep plain launches (an event record between two kernels breaks the chain anyway).
    static const bool pdl_on = !(getenv("GSDF_PDL") != nullptr && getenv("GSDF_PDL")[0] == '0');  // default on; GSDF_PDL=0 is the A/B switch
    auto enqueue = [&](bool stage_events, uint32_t epoch) -> int {
    int rc = 0;
    const bool pdl = pdl_on && m->pdl_chain && !stage_events && !scan3 && !(m->flags & GSDF_MESH_KEEP_GRID);
    if (m->flags & GSDF_MESH_KEEP_GRID) CU(cudaMemsetAsync(m->d_grid, 0x7f, (size_t)D.pitch * (D.ny + 1) * nk * sizeof(float), st));
    // counters and look-back scan state are already zero: re-armed by the previous render's k_finish_render (or by the allocation)
    if (prune) {
        for (int li = 0; li < m->nl3; li++) {  // coarse to fine; each level looks only at the children of kept cubes
            if (m->fine && li == m->nl3 - 1) {  // levels 3 and 2 in one launch
                PruneFine pf{};
                pf.ox = lat.origin[0]; pf.oy = lat.origin[1]; pf.oz = lat.origin[2]; pf.res = lat.res;
                pf.L = m->lev[li];
                if (li) { pf.P = m->lev[li - 1]; pf.shift = m->plan.level[li - 1] - m->plan.level[li]; }
                pf.nx = D.nx; pf.ny = D.ny; pf.nz = D.nz;
                pf.half2 = m->half2; pf.maxDist2 = m->maxDist2;
                pf.bits2 = m->d_bits2; pf.childmask = m->d_childmask;
                pf.kept = m->d_ctr + 4; pf.kept2 = m->d_ctr + 19; pf.evals = m->d_ctr + 7;
                if ((rc = launch_prune_fine(p, pf, st, pdl && li > 0, m->d_ctr + 8 + 2 * li, li == 0 ? m->d_stamp + 0 : nullptr))) return rc;
                continue;
            }
            GenCenters gc{};
            gc.ox = lat.origin[0]; gc.oy = lat.origin[1]; gc.oz = lat.origin[2]; gc.res = lat.res;
            gc.L = m->lev[li];
            if (li) { gc.P = m->lev[li - 1]; gc.shift = m->plan.level[li - 1] - m->plan.level[li]; }
            gc.kept = li == m->nl3 - 1 ? m->d_ctr + 4 : nullptr;
            gc.evals = m->d_ctr + 7;
            if ((rc = launch_centers(p, gc, (uint64_t)gc.L.nwx * 32u * gc.L.ncy * gc.L.ncz, st, pdl && li > 0, m->d_ctr + 8 + 2 * li,
                                     li == 0 ? m->d_stamp + 0 : nullptr))) return rc;
        }
        const uint64_t ncrows = (uint64_t)(D.ny + 1) * nk;
        if (blockmc)
            CU(launch_chain(pdl, k_mesh_lists, dim3(grid_for(p->sms, (ncrows * (uint64_t)((D.nqx + 31) >> 5) + 31) / 32 + ((uint64_t)D.nbz * D.nby * D.nwx + 31) / 32, kThreads / 32)), dim3(kThreads), 0, st, D,
                            (const uint32_t *)m->d_mbits, m->d_list, m->d_ctr + 0, m->d_blklist, m->d_ctr + 5, m->d_stamp + 1,
                            (const uint32_t *)(m->fine ? m->d_bits2 : nullptr), half ? 1 : 0));
        else
        CU(launch_chain(pdl, k_compact_quads, dim3(grid_for(p->sms, (ncrows * (uint64_t)((D.nqx + 31) >> 5) + 31) / 32, kThreads / 32)), dim3(kThreads), 0, st, D, (const uint32_t *)m->d_mbits, m->d_list, m->d_ctr + 0, m->d_stamp + 1));
        CU(cudaGetLastError());
    }
    if (!prune && blockmc) {  // FlatRenderer: every block of the slab is listed (no bit rows, no quad list)
        CU(launch_chain(false, k_mesh_lists, dim3(grid_for(p->sms, ((uint64_t)D.nbz * D.nby * D.nwx + 31) / 32, kThreads / 32)), dim3(kThreads), 0, st, D,
                        (const uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, m->d_blklist, m->d_ctr + 5, (unsigned long long *)nullptr,
                        (const uint32_t *)nullptr, 0));
        CU(cudaGetLastError());
    }
    if (stage_events) CU(cudaEventRecord(m->ev[1], st));
    {
        // with a device-side list length the launch is sized for the worst case; surplus CTAs find no tile and exit
        uint32_t *sched = m->d_ctr + 8 + 2 * GSDF_PRUNE_MAX_LEVELS;
        if (eval_p == 1) {  // latency-bound amount of listed work: one corner per thread, four times as many (short) tiles
            GenGrid<1> g{make_lat(&lat, D.cz0, D.cz0 + nk, D.pitch, true), m->d_grid, prune ? m->d_list : nullptr, prune ? m->d_ctr + 0 : nullptr};
            if ((rc = launch_grid1(p, g, std::min<uint64_t>(nquads, with_margin(m->quad_hint, 1024)) * 4, st, pdl && prune, sched, m->d_stamp + 2))) return rc;
        } else if (eval_p == 2) {
            GenGrid<2> g{make_lat(&lat, D.cz0, D.cz0 + nk, D.pitch, true), m->d_grid, prune ? m->d_list : nullptr, prune ? m->d_ctr + 0 : nullptr};
            g.L.hq = half ? 1 : 0;  // (the bound below is 2 * quads either way: an upper bound of the half-quads listed)
            const uint64_t bound = (prune && m->runs > 0) ? std::min<uint64_t>(nquads, with_margin(m->quad_hint, 1024)) : nquads;
            if ((rc = launch_grid2(p, g, bound * 2, st, pdl && (prune || blockmc), sched, m->d_stamp + 2))) return rc;
        } else {
            GenGrid<4> g{make_lat(&lat, D.cz0, D.cz0 + nk, D.pitch, true), m->d_grid, prune ? m->d_list : nullptr, prune ? m->d_ctr + 0 : nullptr};
            const uint64_t bound = (prune && m->runs > 0) ? std::min<uint64_t>(nquads, with_margin(m->quad_hint, 1024)) : nquads;
            if ((rc = launch_grid4(p, g, bound, st, pdl && (prune || blockmc), sched, m->d_stamp + 2))) return rc;
        }
    }
    if (stage_events) CU(cudaEventRecord(m->ev[2], st));
    const uint64_t ntiles5 = (uint64_t)((D.nsx + 3) / 4) * ((D.ny + kTileY - 1) / kTileY) * ((D.cz1 - D.cz0 + kT5Layers - 1) / kT5Layers);
    const unsigned grid5 = grid_for(p->sms, ntiles5, 1, 8);
    if (blockmc) {
        if (A.cases) CU(cudaMemsetAsync(A.cases, 0, (size_t)ncells, st));  // parity mode only: cells outside kept blocks have case 0
        BlkArgs Cn = BA;
        Cn.stamp = m->d_stamp + 3;
        CU(launch_chain(pdl && !A.cases, k_mc_blk_count, dim3(blkgrid), dim3(kBlkWarps * 32), 0, st, m->tmapB, Cn));
    } else if (m->use_tma && m->tile5) {
        MCArgs Cn = A;
        Cn.stamp = m->d_stamp + 3;
        CU(launch_chain(pdl, k_mc_count5, dim3(grid5), dim3(256), 0, st, m->tmap5, Cn));
    } else if (m->use_tma) {
        const uint64_t ntiles = (uint64_t)((D.nsx + 3) / 4) * ((D.ny + kTileY - 1) / kTileY) * (D.cz1 - D.cz0);
        static const bool count_v1 = getenv("GSDF_COUNT_V1") != nullptr;  // A/B switch: one cell per lane, no prefetch
        // test knob: cap the grid so that small, oracle-checked lattices run many tiles per CTA through both stencil buffers
        static const unsigned count_grid_cap = getenv("GSDF_COUNT_GRID") ? (unsigned)std::max(1, atoi(getenv("GSDF_COUNT_GRID"))) : 0u;
        unsigned cgrid = grid_for(p->sms, ntiles, 1, 16);
        if (count_grid_cap) cgrid = std::min(cgrid, count_grid_cap);
        MCArgs Cn = A;
        Cn.stamp = m->d_stamp + 3;
        CU(launch_chain(pdl, count_v1 ? k_mc_count_tma : k_mc_count_tma4, dim3(cgrid), dim3(256), 0, st, m->tmap, Cn));
    } else {
        MCArgs Cn = A;
        Cn.stamp = m->d_stamp + 3;
        CU(launch_chain(pdl, k_mc_count, dim3(mcgrid), dim3(kThreads), 0, st, Cn));
    }
    CU(cudaGetLastError());
    if (scan3) {
        k_scan_reduce<<<(unsigned)nscanblocks, kThreads, 0, st>>>(m->d_seg, nseg, m->d_blocksum);
        CU(cudaGetLastError());
        k_scan_blocksums<<<1, 1024, 0, st>>>(m->d_blocksum, (uint32_t)nscanblocks, reinterpret_cast<unsigned long long *>(m->d_ctr + 2));
        CU(cudaGetLastError());
        k_scan_apply<<<(unsigned)nscanblocks, kThreads, 0, st>>>(m->d_seg, nseg, m->d_blocksum);
        CU(cudaGetLastError());
    } else if (blockmc && scan_fused && emitted) {
        // no scan launch: the emit pass takes the scan tiles by ticket before it emits (mc_block.cuh)
    } else if (blockmc) {
        CU(launch_chain(pdl, k_scan_seg, dim3((unsigned)nscantiles), dim3(kThreads), 0, st, D, (const uint32_t *)A.mbits, (const uint2 *)m->d_blkcnt, m->d_seg, (uint32_t)nseg,
                        m->d_scanstate, m->d_ctr + 6, epoch, reinterpret_cast<unsigned long long *>(m->d_ctr + 2), m->d_stamp + 4, BA.tilesum ? 1 : 0));
        CU(cudaGetLastError());
    } else {
        CU(launch_chain(pdl, k_scan_lookback, dim3((unsigned)nscantiles), dim3(kThreads), 0, st, m->d_seg, (uint32_t)nseg, m->d_scanstate, m->d_ctr + 6, epoch,
                        reinterpret_cast<unsigned long long *>(m->d_ctr + 2), m->d_stamp + 4));
        CU(cudaGetLastError());
    }
    if (stage_events) CU(cudaEventRecord(m->ev[3], st));
    if (emitted) {
        MCArgs E = A;
        E.cases = nullptr;
        E.stamp = m->d_stamp + 5;
        if (blockmc) {
            BlkArgs BE = BA;
            BE.cases = nullptr;
            BE.stamp = m->d_stamp + 5;
            BE.fin_ctr = m->d_ctr; BE.fin_hctr = (volatile uint32_t *)m->h_ctr; BE.fin_nctr = kMeshCtr;
            BE.fin_scanstate = m->d_scanstate; BE.fin_nstate = (uint32_t)nscantiles;
            BE.fin_dstamp = m->d_stamp; BE.fin_hstamp = (volatile unsigned long long *)m->h_stamp; BE.fin_nstamp = kMeshStamps;
            BE.fin_done = m->d_ctr + kMeshCtr - 1;
            if (scan_fused) {
                BE.scan_state = m->d_scanstate; BE.scan_ticket = m->d_ctr + 6; BE.scan_done = m->d_ctr + 18; BE.scan_epoch = epoch; BE.scan_nseg = (uint32_t)nseg;
                BE.scan_total = reinterpret_cast<unsigned long long *>(m->d_ctr + 2);
            }
            CU(launch_chain(pdl, k_mc_blk_emit, dim3(blkgrid), dim3(kBlkWarps * 32), 0, st, m->tmapB, BE));
        } else if (m->use_tma && m->tile5) {
            // the emit pass ends the render itself: its last CTA publishes the counters and re-arms the state
            E.fin_ctr = m->d_ctr; E.fin_hctr = (volatile uint32_t *)m->h_ctr; E.fin_nctr = kMeshCtr;
            E.fin_scanstate = m->d_scanstate; E.fin_nstate = (uint32_t)nscantiles;
            E.fin_dstamp = m->d_stamp; E.fin_hstamp = (volatile unsigned long long *)m->h_stamp; E.fin_nstamp = kMeshStamps;
            E.fin_done = m->d_ctr + kMeshCtr - 1;
            CU(launch_chain(pdl, k_mc_emit5, dim3(grid5), dim3(256), 0, st, m->tmap5, E));
        } else CU(launch_chain(pdl, k_mc_emit, dim3(mcgrid), dim3(kThreads), 0, st, E));
        CU(cudaGetLastError());
    }
    if (!(emitted && m->use_tma && (m->tile5 || blockmc))) {   // publish the counters (cudaMallocHost memory is device-mapped under UVA) and re-arm the state for the next render
        const uint32_t nstate = (uint32_t)nscantiles;
        CU(launch_chain(pdl, k_finish_render, dim3((unsigned)std::min<uint64_t>(std::max<uint64_t>((nstate + 255) / 256, 1), 64)), dim3(256), 0, st,
                        m->d_ctr, (volatile uint32_t *)m->h_ctr, kMeshCtr, m->d_scanstate, nstate, m->d_stamp, (volatile unsigned long long *)m->h_stamp, kMeshStamps));
        CU(cudaGetLastError());
    }
    return rc;
    };  // enqueue

    // Graph key: everything the captured launches were built from. Any change (buffer regrowth, another program,
    // gsdf_program_update with a different size) re-captures.
    struct GraphKey {
        const void *ptr[10];
        const void *lbits[GSDF_PRUNE_MAX_LEVELS];
        gsdf_prune_plan plan;
        const void *prog;
        const void *jit;   // run-time compiled kernels in use (gsdf_program_specialize after a capture re-captures)
        size_t tri_cap;
        ProgView pv;
        unsigned flags;
        int ext, tma, eval_p, mc_mode, pdl_chain;
        uint32_t quad_hint, blk_hint;
        const void *blk[4];
    } key;
    std::memset(&key, 0, sizeof key);
    const void *kp[10] = {m->d_grid, nullptr, m->d_mbits, m->d_list, m->d_seg, m->d_seglist, m->d_scanstate, m->d_tris, m->d_cases, m->d_segcases};
    std::memcpy(key.ptr, kp, sizeof kp);
    for (int li = 0; li < GSDF_PRUNE_MAX_LEVELS; li++) key.lbits[li] = m->d_lbits[li];
    key.plan = m->plan; key.prog = p;
    key.jit = (p->jit && p->jit->key == p->skey) ? (const void *)p->jit.get() : nullptr;
    key.tri_cap = m->tri_cap; key.pv = p->pv; key.flags = m->flags; key.ext = p->needs_ext ? 1 : 0; key.tma = (m->use_tma ? 1 : 0) | (m->tile5 ? 2 : 0);
    key.eval_p = eval_p; key.quad_hint = (prune && m->runs > 0) ? m->quad_hint : 0u;
    key.pdl_chain = m->pdl_chain ? 1 : 0;
    key.mc_mode = m->mc_mode; key.blk_hint = (blockmc && m->runs > 0) ? m->blk_hint : 0u;
    key.blk[0] = m->d_blklist; key.blk[1] = m->d_blkcnt; key.blk[2] = m->d_bits2; key.blk[3] = m->d_childmask;
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
    m->pending = true; m->pend_graph = use_graph; m->pend_emitted = emitted; m->pendA = A; m->pend_mcgrid = mcgrid; m->pendB = BA; m->pend_blkgrid = blkgrid; m->pend_half = half;
    return 0;
}

// Waits for the enqueued render, reads its counters, re-emits if the triangle buffer was too small, fills the statistics.
int mesh_run_end(gsdf_mesher *m) {
    if (!m->pending) return 0;
    m->pending = false;
    gsdf_program *p = m->prog;
    if (!p) return fail(GSDF_EINVAL, "the renderer's program was destroyed");
    CU(use_device(p->device));
    cudaStream_t st = m->stream;
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
    m->last_reemit = !emitted || total * 9 > m->tri_cap;
    if (!emitted || total * 9 > m->tri_cap) {
        if (m->copy_stream) CU(cudaStreamSynchronize(m->copy_stream));  // a speculative prefix read may be using d_tris
        if ((rc = grow(m->d_tris, m->tri_cap, (size_t)std::max<uint64_t>(total, 1) * 9))) return rc;
        A.tris = m->d_tris;
        A.tri_capacity = m->tri_cap / 9;
        A.cases = nullptr;
        // k_finish_render re-armed the counters already: give the emit its segment-list length back, clear again after
        CU(cudaMemcpyAsync(m->d_ctr + 5, m->h_ctr + 5, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        if (m->use_tma && m->mc_mode == 2 && prune) {
            BlkArgs BE = m->pendB;
            BE.tris = m->d_tris; BE.tri_capacity = m->tri_cap / 9; BE.cases = nullptr; BE.stamp = nullptr;
            k_mc_blk_emit<<<m->pend_blkgrid, kBlkWarps * 32, 0, st>>>(m->tmapB, BE);
        } else if (m->use_tma && m->tile5) {
            const uint64_t ntiles5 = (uint64_t)((D.nsx + 3) / 4) * ((D.ny + kTileY - 1) / kTileY) * ((D.cz1 - D.cz0 + kT5Layers - 1) / kT5Layers);
            A.stamp = nullptr;
            k_mc_emit5<<<grid_for(p->sms, ntiles5, 1, 8), 256, 0, st>>>(m->tmap5, A);
        } else {
            k_mc_emit<<<mcgrid, kThreads, 0, st>>>(A);
        }
        CU(cudaGetLastError());
        CU(cudaMemsetAsync(m->d_ctr, 0, 8 * sizeof(uint32_t), st));  // (the scheduler pairs behind them reset themselves)
        CU(cudaEventRecord(m->ev[4], st));
        CU(cudaStreamSynchronize(st));
    }
    m->ntri = total;
    m->read_pos = 0;
    if (prune) {
        // prune-cube centres of every level + the listed lattice quads (4 corners each) or half-quads (2 corners each)
        m->evals = (uint64_t)m->h_ctr[7] + (m->pend_half ? 2ull : 4ull) * m->h_ctr[0];
        // hint for the next render's launch shape, in quads; rounded up so that small changes of the tree do not re-capture the graph
        m->quad_hint = ((m->pend_half ? (m->h_ctr[0] + 1u) / 2u : m->h_ctr[0]) + 4095u) & ~4095u;
    }
    m->blk_hint = (m->h_ctr[5] + 1023u) & ~1023u;  // listed blocks (block kernels) of this render
    if (prune && m->fine) {  // the finest cubes of the plan are 2 cells wide: Cube.DecomposesTo(1) = 8
        const uint64_t n2 = (uint64_t)((D.nx + 1) / 2) * ((D.ny + 1) / 2) * (uint64_t)(((D.cz1 + 1) >> 1) - (D.cz0 >> 1));
        m->pruned = (n2 - m->h_ctr[19]) * 8ull;
    } else if (prune) {
        m->pruned = (nblocks - m->h_ctr[4]) * 64ull;  // Cube.DecomposesTo(1) of a level-3 cube = 8^2
    } else {
        m->evals = (uint64_t)(D.nx + 1) * (D.ny + 1) * nk;
        m->pruned = 0;
    }
    if (use_graph) {
        // no events between the kernels of a graph replay: the stage times come from the kernels' own %globaltimer stamps
        const unsigned long long *t = m->h_stamp;
        const unsigned long long t0 = prune ? t[0] : t[2];
        auto msd = [](unsigned long long a, unsigned long long b) { return b > a ? (float)((double)(b - a) * 1e-6) : 0.f; };
        m->ms[0] = msd(t0, t[2]); m->ms[1] = msd(t[2], t[3]); m->ms[2] = msd(t[3], t[5]); m->ms[3] = msd(t[5], t[kMeshStamps - 1]);
    } else { for (int i = 0; i < 4; i++) cudaEventElapsedTime(&m->ms[i], m->ev[i], m->ev[i + 1]); }
    cudaEventElapsedTime(&m->ms[4], m->ev[0], m->ev[4]);
    m->runs++;
    p->evals += m->evals;  // the reference's renderers evaluate through sdf.Evaluate: its counter includes them (gleval/gpu.go:80)
    return 0;
}

int mesh_run(gsdf_mesher *m) {
    int rc = mesh_run_begin(m);
    return rc ? rc : mesh_run_end(m);
}

}  // namespace

extern "C" {

int gsdf_prune_plan_default(const gsdf_lattice *lat, unsigned flags, gsdf_prune_plan *out) {
    if (!lat || !out) return fail(GSDF_EINVAL, "gsdf_prune_plan_default: NULL argument");
    if (lat->n[0] <= 0 || lat->n[1] <= 0 || lat->n[2] <= 0) return fail(GSDF_ERES, "resolution not fine enough for marching cubes");
    *out = gsdf_prune_plan{};
    const uint64_t nblocks = (uint64_t)((lat->n[0] + 3) / 4) * ((lat->n[1] + 3) / 4) * ((lat->n[2] + 3) / 4);
    int n = 0;
    // Coarse levels only pay on large lattices: on a small one their launch costs more than the centres they save.
    if (nblocks >= (64ull << 20)) { out->level[n] = 7; out->margin[n++] = GSDF_PRUNE_MARGIN_DEFAULT; }
    if (nblocks >= (1ull << 20)) { out->level[n] = 5; out->margin[n++] = GSDF_PRUNE_MARGIN_DEFAULT; }
    out->level[n] = 3;
    out->margin[n++] = (flags & GSDF_MESH_PRUNE_LITERAL) ? 1.0f : GSDF_PRUNE_MARGIN_DEFAULT;
    // The 2-cell level (corners of dropped 2-cell cubes inside kept 4-cell blocks are not evaluated: -37 % evaluations on the
    // reference's example parts) is NOT part of the default plan: at margin 1.25 it loses 714 of the 309,872 triangles of the
    // README's fibonacci-showerhead run (a steep, non-Lipschitz field; margin 2 would be needed and keeps 82 % of the cubes),
    // and on flange@400 its extra rounds cost the prune stage what the evaluation saves. Explicit plans may end
    // with it; GSDF_PRUNE_FINE=1 adds it to the default plan (A/B switch).
    static const bool fine_on = getenv("GSDF_PRUNE_FINE") != nullptr && getenv("GSDF_PRUNE_FINE")[0] == '1';
    // (the 2-cell level lives in the kept-block marching-cubes kernels: not with the A/B kernel sets)
    const char *mc = getenv("GSDF_MC");
    const bool blockmode = !getenv("GSDF_NO_TMA") && !getenv("GSDF_MC_TILE5") && (!mc || (strcmp(mc, "v1") != 0 && strcmp(mc, "tile5") != 0));
    if (fine_on && blockmode && !(flags & GSDF_MESH_PRUNE_LITERAL)) { out->level[n] = 2; out->margin[n++] = GSDF_PRUNE_MARGIN_DEFAULT; }
    out->nlevels = n;
    return 0;
}

int gsdf_mesh_begin(gsdf_program *p, const gsdf_lattice *lat, int cz0, int cz1, unsigned flags, gsdf_mesher **out) {
    return gsdf_mesh_begin_plan(p, lat, cz0, cz1, flags & ~(unsigned)GSDF_MESH_PLAN_GIVEN, nullptr, out);
}

static int mesh_begin_prio(gsdf_program *p, const gsdf_lattice *lat, int cz0, int cz1, unsigned flags, const gsdf_prune_plan *plan,
                           int stream_priority, gsdf_mesher **out);

int gsdf_mesh_begin_plan(gsdf_program *p, const gsdf_lattice *lat, int cz0, int cz1, unsigned flags, const gsdf_prune_plan *plan,
                         gsdf_mesher **out) {
    return mesh_begin_prio(p, lat, cz0, cz1, flags, plan, 0, out);
}

// stream_priority: 0 = default; -k = k steps towards the device's highest stream priority (the multi-device mesher gives the
// slabs whose triangles leave first the right of way, so that their read-back starts while the later slabs still compute)
static int mesh_begin_prio(gsdf_program *p, const gsdf_lattice *lat, int cz0, int cz1, unsigned flags, const gsdf_prune_plan *plan,
                           int stream_priority, gsdf_mesher **out) {
    if (!p || !lat || !out) return fail(GSDF_EINVAL, "gsdf_mesh_begin: NULL argument");
    if (p->dim != 3) return fail(GSDF_EINVAL, "program is not 3D");
    if (!(lat->res > 0) || lat->n[0] <= 0 || lat->n[1] <= 0 || lat->n[2] <= 0) return fail(GSDF_ERES, "resolution not fine enough for marching cubes");
    if (cz0 < 0 || cz1 > lat->n[2] || cz0 >= cz1) return fail(GSDF_EINVAL, "bad cell slab [%d,%d)", cz0, cz1);
    gsdf_prune_plan pl{};
    if (plan) {
        flags |= GSDF_MESH_PRUNE;
        pl = *plan;
        const bool ends2 = pl.nlevels >= 2 && pl.nlevels <= GSDF_PRUNE_MAX_LEVELS && pl.level[pl.nlevels - 1] == 2 && pl.level[pl.nlevels - 2] == 3;
        if (pl.nlevels < 1 || pl.nlevels > GSDF_PRUNE_MAX_LEVELS || (pl.level[pl.nlevels - 1] != 3 && !ends2))
            return fail(GSDF_EINVAL, "prune plan: 1..%d levels ending with level 3, or with level 3 followed by level 2", GSDF_PRUNE_MAX_LEVELS);
        for (int i = 0; i < pl.nlevels; i++) {
            if (pl.level[i] < 2 || pl.level[i] > 12 || (i && pl.level[i] >= pl.level[i - 1])) return fail(GSDF_EINVAL, "prune plan: levels must descend strictly within [2, 12]");
            if (!(pl.margin[i] >= 1.0f) || std::isinf(pl.margin[i])) return fail(GSDF_EINVAL, "prune plan: margins must be finite and >= 1");
        }
    } else if (flags & GSDF_MESH_PRUNE) {
        int prc = gsdf_prune_plan_default(lat, flags, &pl);
        if (prc) return prc;
    }
    int rc = ensure_device(p->device);
    if (rc) return rc;
    gsdf_mesher *m = new gsdf_mesher();
    m->prog = p;
    m->device = p->device;
    m->lat = *lat;
    m->flags = flags;
    m->plan = pl;
    m->use_tma = getenv("GSDF_NO_TMA") == nullptr;  // A/B switch for the classification kernel
    // A/B switch. GSDF_MC_TILE5=1 selects the 4-layer-tile pair (mc_tile5.cuh): measured on B200 (round 2) its count pass is
    // 6 us faster than the one-layer-tile pass, its emit pass 25 us slower than the work-list emit (it classifies again, and
    // a warp walks its row's segments serially) -- kept for the record, not the default.
    m->tile5 = getenv("GSDF_MC_TILE5") != nullptr;
    if (const char *mc = getenv("GSDF_MC")) {
        if (!strcmp(mc, "v1")) m->mc_mode = 0;
        else if (!strcmp(mc, "tile5")) { m->mc_mode = 1; m->tile5 = true; }
        else m->mc_mode = 2;
    } else if (m->tile5) m->mc_mode = 1;
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
    fastdiv_init((uint32_t)D.nbx, D.nbx_mul, D.nbx_shr);
    fastdiv_init((uint32_t)D.nby, D.nby_mul, D.nby_shr);
    m->fine = pl.nlevels >= 2 && pl.level[pl.nlevels - 1] == 2;
    // the 2-cell level lives in the kept-block marching-cubes kernels only; the A/B kernel sets fall back to the plan's level 3
    if (m->fine && !(m->use_tma && m->mc_mode == 2)) { m->fine = false; pl.nlevels--; m->plan = pl; }
    m->nl3 = pl.nlevels - (m->fine ? 1 : 0);
    if (m->fine) {
        const float size2 = lat->res * 2.0f;
        m->half2 = size2 * 0.5f;
        m->maxDist2 = size2 * (float)(1.73205080757 / 2) * pl.margin[pl.nlevels - 1];
    }
    for (int li = 0; li < m->nl3; li++) {  // cubes of level L are 2^(L-1) cells wide and aligned to the lattice origin
        PruneLevel &Lv = m->lev[li];
        Lv.w = 1 << (pl.level[li] - 1);
        Lv.ncx = (D.nx + Lv.w - 1) / Lv.w; Lv.ncy = (D.ny + Lv.w - 1) / Lv.w;
        Lv.cz0 = cz0 / Lv.w;
        Lv.ncz = (cz1 + Lv.w - 1) / Lv.w - Lv.cz0;
        Lv.nwx = (Lv.ncx + 31) / 32;
        Lv.fdiv = (uint64_t)Lv.nwx * 32u * (uint64_t)Lv.ncy * (uint64_t)std::max(Lv.ncz, 1) < (1ull << 31) ? 1u : 0u;
        fastdiv_init((uint32_t)Lv.nwx * 32u, Lv.row_mul, Lv.row_shr);
        fastdiv_init((uint32_t)Lv.ncy, Lv.ncy_mul, Lv.ncy_shr);
        const float size = lat->res * (float)Lv.w;                            // ms3.Octree.CubeSize
        Lv.half = size * 0.5f;
        Lv.maxDist = size * (float)(1.73205080757 / 2) * pl.margin[li];         // octreerenderer.go:182 with glrender.go:9, times the margin
    }
    cudaError_t e = cudaMalloc((void **)&m->d_ctr, kMeshCtr * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(m->d_ctr, 0, kMeshCtr * sizeof(uint32_t));  // every render leaves them zeroed for the next (k_finish_render)
    if (e == cudaSuccess) e = cudaMallocHost((void **)&m->h_ctr, kMeshCtr * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_stamp, kMeshStamps * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(m->d_stamp, 0, kMeshStamps * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&m->h_stamp, kMeshStamps * sizeof(unsigned long long));
    for (int i = 0; i < 5 && e == cudaSuccess; i++) e = cudaEventCreate(&m->ev[i]);
    if (e == cudaSuccess) {
        int least = 0, greatest = 0;
        e = cudaDeviceGetStreamPriorityRange(&least, &greatest);  // numerically lower = higher priority
        const int prio = std::max(greatest, std::min(least, least + stream_priority));
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&m->stream, cudaStreamNonBlocking, prio);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { gsdf_mesh_destroy(m); return fail(GSDF_ECUDA, "mesher setup: %s", cudaGetErrorString(e)); }
    program_add_dependent(p, m->ev[4], &m->prog);
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
    CU(use_device(m->device));
    const uint64_t n = std::min<uint64_t>(ntris, m->tri_cap / 9);
    if (n == 0) return 0;
    CU(cudaStreamWaitEvent(m->copy_stream, m->ev[4], 0));  // after the emit of the render enqueued last
    CU(cudaMemcpyAsync(tri9, m->d_tris, n * 9 * sizeof(float), cudaMemcpyDeviceToHost, m->copy_stream));
    return (int64_t)n;
}

int gsdf_mesh_set_program(gsdf_mesher *m, gsdf_program *p) {
    if (!m || !p) return fail(GSDF_EINVAL, "gsdf_mesh_set_program: NULL argument");
    if (p->dim != 3) return fail(GSDF_EINVAL, "program is not 3D");
    if (p->device != m->device) return fail(GSDF_EINVAL, "program lives on another device");
    if (m->pending) { const int erc = mesh_run_end(m); if (erc) return erc; }
    program_remove_dependent(m->prog, m->ev[4]);
    program_add_dependent(p, m->ev[4], &m->prog);
    m->prog = p;
    return 0;
}

int64_t gsdf_mesh_read(gsdf_mesher *m, float *tri9, size_t max_tris) {
    if (!m || !tri9) return fail(GSDF_EINVAL, "gsdf_mesh_read: NULL argument");
    if (m->pending) { const int erc = mesh_run_end(m); if (erc) return erc; }
    if (max_tris < 5) return fail(GSDF_ESHORT, "short buffer");  // flatrenderer.go:187
    CU(use_device(m->device));
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
    CU(use_device(m->device));
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
    CU(use_device(m->device));
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
    CU(use_device(m->device));
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
    CU(use_device(m->device));
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
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->prog && m->ev[4]) program_remove_dependent(m->prog, m->ev[4]);
    for (auto &b : m->d_lbits) cudaFree(b);
    cudaFree(m->d_bits2); cudaFree(m->d_childmask);
    cudaFree(m->d_grid); cudaFree(m->d_list); cudaFree(m->d_seg); cudaFree(m->d_seglist); cudaFree(m->d_segcases); cudaFree(m->d_scanstate); cudaFree(m->d_blocksum);
    cudaFree(m->d_blklist); cudaFree(m->d_blkcnt);
    cudaFree(m->d_tris); cudaFree(m->d_cases); cudaFree(m->d_stl); cudaFree(m->d_ctr);
    if (m->h_ctr) cudaFreeHost(m->h_ctr);
    cudaFree(m->d_stamp);
    if (m->h_stamp) cudaFreeHost(m->h_stamp);
    for (auto &e : m->ev) if (e) cudaEventDestroy(e);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->stream) cudaStreamDestroy(m->stream);
    if (m->gexec) cudaGraphExecDestroy(m->gexec);
    delete m;
    (void)cudaGetLastError();  // teardown never leaves a stale (non-sticky) error behind for the next launch check
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ multi-device mesher
// One lattice = nslabs Z-slabs dealt round-robin over ndev devices (slab j on device j % ndev). Worker w (one host thread per
// device; worker 0 is the calling thread) owns the slabs j = w, w + ndev, ... Each render:
//   1. every worker uploads a pending program update to its device and enqueues ALL its slabs (each slab is a gsdf_mesher
//      with its own stream and CUDA graph, so the slabs of one device overlap on the GPU);
//   2. in slab order it waits for a slab's counters, publishes the triangle count, waits until the counts of all earlier
//      slabs (other workers') are published -- that sum is the slab's offset in the caller's buffer -- and enqueues the
//      device->host copy on the slab's copy stream, which runs under the kernels of the later slabs;
//   3. it waits for its copies.
// No collective and no speculation: offsets come from counts actually read.
struct gsdf_multimesher {
    int ndev = 0, nslabs = 0;
    std::vector<int> devs;
    std::vector<gsdf_program *> prog;    // per device
    std::vector<gsdf_mesher *> slab;     // per slab (slab j on devs[j % ndev])
    std::vector<int32_t> cuts;           // nslabs + 1
    gsdf_lattice lat{};
    unsigned flags = 0;
    // pending program update (applied by every worker at the start of the next render)
    std::vector<uint8_t> up_blob;
    std::vector<float> up_aux;
    std::vector<uint8_t> up_dirty;       // per device
    // pinned staging per device for destinations that are not page-locked
    std::vector<float *> h_stage;
    std::vector<size_t> h_stage_cap;
    // job hand-off (spin first, then sleep: renders last a few hundred microseconds)
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> job_seq{0};
    std::atomic<int> job_done{0};
    std::atomic<bool> quit{false};
    // state of the render in flight
    float *dst = nullptr;
    size_t max_tris = 0;
    bool dst_pinned = false;
    std::vector<std::atomic<int64_t>> count;   // per slab: -1 = not yet known
    std::atomic<int> abort_rc{0};
    std::vector<std::string> worker_err;       // per worker: message of its failure
    // results
    uint64_t ntri = 0, evals = 0, pruned = 0, read_pos = 0;
    std::vector<uint64_t> offs;                // per slab triangle offset of the last render
    float device_ms = 0;
    // host-clock timeline of the last render in microseconds from the call: [0] slabs enqueued (worker 0), then per slab
    // {count seen, copy enqueued}, last = everything delivered
    std::vector<double> timeline;
    std::chrono::steady_clock::time_point t_call;
    bool rendered = false;
    bool delivered = false;                    // the last render's triangles are already in the caller's buffer
    cudaStream_t copyk_stream = nullptr;       // device-driven read-back (k_copy_out), one device only
    explicit gsdf_multimesher(int n) : count(n) {}
};

static int mesh_begin_prio(gsdf_program *p, const gsdf_lattice *lat, int cz0, int cz1, unsigned flags, const gsdf_prune_plan *plan,
                           int stream_priority, gsdf_mesher **out);

namespace {

// Programmatic dependent launch per slab. The chain of the slab that leaves first is latency critical and keeps it; under
// it the CTAs of kernel i+1 are resident (waiting) while kernel i runs, which on a device shared by several slabs keeps the
// other slabs' kernels out (measured: three slabs ran one after the other, 68 + 106 + 98 us). The later slabs of a device
// are launched plainly and overlap. GSDF_MULTI_PDL=all|first|none is the A/B switch (default first).
bool multi_slab_pdl(const gsdf_multimesher *mm, int j) {
    static const char *mode = getenv("GSDF_MULTI_PDL");
    if (mode && !strcmp(mode, "all")) return true;
    if (mode && !strcmp(mode, "none")) return false;
    return j / mm->ndev == 0;
}

// the k-th slab of a device leaves k-th: earlier slabs get the higher stream priority
int multi_slab_priority(const gsdf_multimesher *mm, int j) {
    const int nper = (mm->nslabs + mm->ndev - 1) / mm->ndev;
    return -(nper - 1 - j / mm->ndev);
}

// Device-driven read-back (A/B: GSDF_MULTI_COPYK=1, one device, page-locked destination). The host normally learns a slab's
// triangle count from the counters its last kernel publishes and then enqueues the copy -- and while an earlier slab's DMA is
// in flight that news arrives ~50 us late. Here the copy of slab j is a small kernel enqueued up front on one
// high-priority stream behind the slab's completion event: it reads the counts of slabs 0..j from their mapped host counters
// itself, and writes the slab's triangles into the caller's buffer at the right offset in 16-byte aligned vectors. CTAs of 128
// threads x <= 32 registers fit beside the persistent grids of the slabs still running (they all leave 4096 registers per SM).
constexpr int kCopyKMaxSlabs = 32;
struct CopyOutArgs {
    const float *src;
    float *dst;
    const uint32_t *hctr[kCopyKMaxSlabs];
    int j;
    unsigned long long max_tris, src_cap_tris;
};
constexpr int kCopyKThreads = 128;
__global__ void __launch_bounds__(kCopyKThreads, 16) k_copy_out(const CopyOutArgs a) {
    __shared__ unsigned long long s_off, s_n;
    if (threadIdx.x == 0) {
        unsigned long long off = 0;
        for (int i = 0; i < a.j; i++) off += *reinterpret_cast<const volatile unsigned long long *>(a.hctr[i] + 2);
        s_off = off;
        s_n = *reinterpret_cast<const volatile unsigned long long *>(a.hctr[a.j] + 2);
    }
    __syncthreads();
    const unsigned long long off = s_off, n = s_n;
    // (count beyond the buffer the slab emitted into, or beyond the destination: the host notices the same and copies the classic way)
    if (n == 0 || n > a.src_cap_tris || off + n > a.max_tris || n >= (1ull << 28)) return;
    const uint32_t nfl = (uint32_t)n * 9u;
    float *d = a.dst + off * 9ull;
    const float *s = a.src;
    uint32_t head = ((16u - (uint32_t)(reinterpret_cast<uintptr_t>(d) & 15u)) & 15u) / 4u;
    if (head > nfl) head = nfl;
    const uint32_t nv = (nfl - head) / 4u, tail = nfl - head - 4u * nv;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (tid < head) d[tid] = s[tid];
    if (tid < tail) d[head + 4u * nv + tid] = s[head + 4u * nv + tid];
    float4 *dv = reinterpret_cast<float4 *>(d + head);
    const float *sv = s + head;
    for (uint32_t i = tid; i < nv; i += 2u * nth) {
        const uint32_t k1 = i + nth;
        const float4 v0 = make_float4(__ldcs(sv + 4u * i), __ldcs(sv + 4u * i + 1), __ldcs(sv + 4u * i + 2), __ldcs(sv + 4u * i + 3));
        float4 v1 = v0;
        if (k1 < nv) v1 = make_float4(__ldcs(sv + 4u * k1), __ldcs(sv + 4u * k1 + 1), __ldcs(sv + 4u * k1 + 2), __ldcs(sv + 4u * k1 + 3));
        dv[i] = v0;
        if (k1 < nv) dv[k1] = v1;
    }
}

// one device, page-locked destination: every slab's read-back enqueued up front (see k_copy_out)
int multi_render_copyk(gsdf_multimesher *mm) {
    int rc;
    auto now_us = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - mm->t_call).count(); };
    if (!mm->copyk_stream) {
        int least = 0, greatest = 0;
        CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CU(cudaStreamCreateWithPriority(&mm->copyk_stream, cudaStreamNonBlocking, greatest));
    }
    static const int ctas = getenv("GSDF_MULTI_COPYK_CTAS") ? std::max(1, atoi(getenv("GSDF_MULTI_COPYK_CTAS"))) : 148;
    for (int j = 0; j < mm->nslabs; j++)
        if ((rc = mesh_run_begin(mm->slab[j]))) return rc;
    for (int j = 0; j < mm->nslabs; j++) {
        gsdf_mesher *m = mm->slab[j];
        CopyOutArgs a{};
        a.src = m->d_tris; a.dst = mm->dst; a.j = j; a.max_tris = mm->max_tris; a.src_cap_tris = m->tri_cap / 9;
        for (int i = 0; i <= j; i++) a.hctr[i] = mm->slab[i]->h_ctr;
        CU(cudaStreamWaitEvent(mm->copyk_stream, m->ev[4], 0));
        k_copy_out<<<ctas, kCopyKThreads, 0, mm->copyk_stream>>>(a);
        CU(cudaGetLastError());
    }
    mm->timeline[0] = now_us();
    uint64_t off = 0;
    bool redo_from_here = false;
    for (int j = 0; j < mm->nslabs; j++) {
        gsdf_mesher *m = mm->slab[j];
        if ((rc = mesh_run_end(m))) return rc;
        mm->count[j].store((int64_t)m->ntri, std::memory_order_release);
        mm->timeline[1 + 2 * j] = mm->timeline[2 + 2 * j] = now_us();
        mm->offs[j] = off;
        // a slab that had to grow its buffer and emit again was skipped by its copy kernel (count > capacity): classic copy
        if ((m->last_reemit || m->ntri >= (1ull << 28)) && m->ntri && off + m->ntri <= mm->max_tris) {
            CU(cudaStreamWaitEvent(m->copy_stream, m->ev[4], 0));
            CU(cudaMemcpyAsync(mm->dst + 9 * off, m->d_tris, m->ntri * 9 * sizeof(float), cudaMemcpyDeviceToHost, m->copy_stream));
            redo_from_here = true;
        }
        off += m->ntri;
    }
    CU(cudaStreamSynchronize(mm->copyk_stream));
    if (redo_from_here)
        for (int j = 0; j < mm->nslabs; j++) CU(cudaStreamSynchronize(mm->slab[j]->copy_stream));
    return 0;
}

int multi_worker_render(gsdf_multimesher *mm, int w) {
    const int dev = mm->devs[w];
    CU(use_device(dev));
    int rc;
    if (mm->up_dirty[w]) {  // no host synchronisation: the slabs' streams wait for the upload on the device
        if ((rc = program_update_async(mm->prog[w], mm->up_blob.data(), mm->up_blob.size(), mm->up_aux.data(), mm->up_aux.size()))) return rc;
        mm->up_dirty[w] = 0;
    }
    auto now_us = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - mm->t_call).count(); };
    static const bool copyk = getenv("GSDF_MULTI_COPYK") != nullptr && getenv("GSDF_MULTI_COPYK")[0] == '1';  // A/B switch
    if (copyk && mm->ndev == 1 && mm->dst && mm->dst_pinned && mm->nslabs <= kCopyKMaxSlabs) return multi_render_copyk(mm);
    for (int j = w; j < mm->nslabs; j += mm->ndev)
        if ((rc = mesh_run_begin(mm->slab[j]))) return rc;
    if (w == 0) mm->timeline[0] = now_us();
    for (int j = w; j < mm->nslabs; j += mm->ndev) {
        gsdf_mesher *m = mm->slab[j];
        if ((rc = mesh_run_end(m))) return rc;
        mm->count[j].store((int64_t)m->ntri, std::memory_order_release);
        mm->timeline[1 + 2 * j] = now_us();
        if (!mm->dst) continue;
        uint64_t off = 0;
        for (int i = 0; i < j; i++) {
            int64_t c;
            while ((c = mm->count[i].load(std::memory_order_acquire)) < 0) {
                if (mm->abort_rc.load(std::memory_order_relaxed)) return 0;  // another worker failed: it reports
                std::this_thread::yield();
            }
            off += (uint64_t)c;
        }
        mm->offs[j] = off;
        if (off + m->ntri > mm->max_tris) continue;  // too small: the caller gets GSDF_ESHORT with nothing guaranteed
        if (m->ntri == 0) continue;
        if (mm->dst_pinned) {
            CU(cudaStreamWaitEvent(m->copy_stream, m->ev[4], 0));
            CU(cudaMemcpyAsync(mm->dst + 9 * off, m->d_tris, m->ntri * 9 * sizeof(float), cudaMemcpyDeviceToHost, m->copy_stream));
            mm->timeline[2 + 2 * j] = now_us();
        } else {  // pageable destination: DMA into this device's pinned staging, then one host copy
            const size_t need = (size_t)m->ntri * 9;
            if (mm->h_stage_cap[w] < need) {
                if (mm->h_stage[w]) cudaFreeHost(mm->h_stage[w]);
                mm->h_stage[w] = nullptr; mm->h_stage_cap[w] = 0;
                if (cudaHostAlloc((void **)&mm->h_stage[w], (need + need / 8) * sizeof(float), cudaHostAllocPortable) != cudaSuccess)
                    return fail(GSDF_ENOMEM, "pinned staging for %zu triangles", (size_t)m->ntri);
                mm->h_stage_cap[w] = need + need / 8;
            }
            CU(cudaStreamWaitEvent(m->copy_stream, m->ev[4], 0));
            CU(cudaMemcpyAsync(mm->h_stage[w], m->d_tris, need * sizeof(float), cudaMemcpyDeviceToHost, m->copy_stream));
            CU(cudaStreamSynchronize(m->copy_stream));
            std::memcpy(mm->dst + 9 * off, mm->h_stage[w], need * sizeof(float));
        }
    }
    if (mm->dst && mm->dst_pinned)
        for (int j = w; j < mm->nslabs; j += mm->ndev) CU(cudaStreamSynchronize(mm->slab[j]->copy_stream));
    return 0;
}

void multi_worker_job(gsdf_multimesher *mm, int w) {
    const int rc = multi_worker_render(mm, w);
    if (rc) {
        mm->worker_err[w] = gsdf_last_error();
        int expect = 0;
        mm->abort_rc.compare_exchange_strong(expect, rc);
        // publish something for every slab of this worker so that nobody spins forever
        for (int j = w; j < mm->nslabs; j += mm->ndev) {
            int64_t c = -1;
            mm->count[j].compare_exchange_strong(c, 0);
        }
    }
    mm->job_done.fetch_add(1, std::memory_order_release);
}

void multi_worker_main(gsdf_multimesher *mm, int w) {
    uint64_t seen = 0;
    for (;;) {
        // spin briefly (the next render usually follows at once), then sleep
        bool got = false;
        for (int spin = 0; spin < 20000; spin++) {
            if (mm->quit.load(std::memory_order_acquire)) return;
            if (mm->job_seq.load(std::memory_order_acquire) != seen) { got = true; break; }
            std::this_thread::yield();
        }
        if (!got) {
            std::unique_lock<std::mutex> lk(mm->mu);
            mm->cv.wait(lk, [&] { return mm->quit.load() || mm->job_seq.load() != seen; });
            if (mm->quit.load()) return;
        }
        seen = mm->job_seq.load(std::memory_order_acquire);
        multi_worker_job(mm, w);
    }
}

// gsdf_slab_cuts: near-equal slabs; interior cuts aligned down to the 4-layer prune blocks while the slabs are at least two
// blocks thick (thinner slabs keep the exact split: evaluating a block's centre twice is cheaper than an empty slab)
void slab_cuts(int nz, int nslabs, int32_t *cuts) {
    const bool align = nz / nslabs >= 8;
    cuts[0] = 0;
    for (int g = 1; g < nslabs; g++) {
        int c = (int)(((int64_t)g * nz) / nslabs);
        if (align) c = (c / 4) * 4;
        cuts[g] = std::max(c, cuts[g - 1]);
    }
    cuts[nslabs] = nz;
}

// New cuts that equalise the cost per slab, from the cost each current slab reported: the cost is spread evenly over the
// slab's layers (piecewise-constant density) and cut g lands where the cumulative cost reaches g/nslabs of the total. Every
// slab keeps at least one layer. Cuts are not block-aligned: a straddled prune block has its centre evaluated by both
// neighbours, which is cheaper than the imbalance a 4-layer granularity leaves on thin lattices.
void slab_rebalance(int nz, int nslabs, const int32_t *cuts, const double *cost, int32_t *out, const double *share = nullptr) {
    double total = 0, wsum = 0;
    for (int j = 0; j < nslabs; j++) { total += std::max(cost[j], 0.0); wsum += share ? share[j] : 1.0; }
    out[0] = 0; out[nslabs] = nz;
    if (!(total > 0) || nslabs < 2) { for (int g = 1; g < nslabs; g++) out[g] = cuts[g]; return; }
    int j = 0;
    double before = 0;  // cost of the slabs in front of slab j
    double wacc = 0;    // share of the new slabs in front of cut g
    for (int g = 1; g < nslabs; g++) {
        wacc += share ? share[g - 1] : 1.0;
        const double target = total * wacc / wsum;
        while (j < nslabs - 1 && before + std::max(cost[j], 0.0) < target) { before += std::max(cost[j], 0.0); j++; }
        const double w = std::max(cost[j], 0.0);
        const double frac = w > 0 ? (target - before) / w : 0.0;
        int c = cuts[j] + (int)std::lround(frac * (cuts[j + 1] - cuts[j]));
        c = std::max(c, out[g - 1] + 1);
        c = std::min(c, nz - (nslabs - g));
        out[g] = c;
    }
}

}  // namespace

extern "C" {

int gsdf_slab_rebalance(int nz, int nslabs, const int32_t *cuts, const double *cost, int32_t *out) {
    if (!cuts || !cost || !out || nslabs < 1 || nz < nslabs) return fail(GSDF_EINVAL, "gsdf_slab_rebalance: bad argument");
    if (cuts[0] != 0 || cuts[nslabs] != nz) return fail(GSDF_EINVAL, "gsdf_slab_rebalance: cuts must run from 0 to nz");
    for (int j = 0; j < nslabs; j++) if (cuts[j + 1] <= cuts[j]) return fail(GSDF_EINVAL, "gsdf_slab_rebalance: empty slab %d", j);
    slab_rebalance(nz, nslabs, cuts, cost, out);
    return 0;
}

int gsdf_slab_cuts(int nz, int nslabs, int32_t *cuts) {
    if (!cuts || nz <= 0 || nslabs <= 0) return fail(GSDF_EINVAL, "gsdf_slab_cuts: nz and nslabs must be positive");
    slab_cuts(nz, nslabs, cuts);
    return 0;
}

int gsdf_multi_begin(int ndev, const int *devs, int slabs_per_device, const void *blob, size_t blob_bytes, const float *aux,
                     size_t aux_floats, const gsdf_lattice *lat, unsigned flags, gsdf_multimesher **out) {
    if (!devs || !lat || !out || ndev < 1 || ndev > 64) return fail(GSDF_EINVAL, "gsdf_multi_begin: bad argument");
    if (slabs_per_device < 1 || slabs_per_device > 64) return fail(GSDF_EINVAL, "gsdf_multi_begin: slabs_per_device must be in [1, 64]");
    if (!(lat->res > 0) || lat->n[0] <= 0 || lat->n[1] <= 0 || lat->n[2] <= 0) return fail(GSDF_ERES, "resolution not fine enough for marching cubes");
    if (flags & (GSDF_MESH_KEEP_CASES | GSDF_MESH_KEEP_GRID | GSDF_MESH_STAGE_TIMING)) return fail(GSDF_EINVAL, "gsdf_multi_begin: parity / timing flags belong to single meshers");
    int ndevs_visible = gsdf_device_count();
    if (ndevs_visible < 0) return ndevs_visible;
    if (ndevs_visible == 0) return fail(GSDF_ECUDA, "no CUDA device available; libgsdfb200 has no CPU fallback");
    for (int i = 0; i < ndev; i++)
        if (devs[i] < 0 || devs[i] >= ndevs_visible) return fail(GSDF_EINVAL, "device %d out of range (have %d)", devs[i], ndevs_visible);
    int nslabs = std::min(ndev * slabs_per_device, (int)lat->n[2]);
    nslabs = std::max(nslabs, 1);
    gsdf_multimesher *mm = new gsdf_multimesher(nslabs);
    mm->ndev = std::min(ndev, nslabs); mm->nslabs = nslabs; mm->lat = *lat; mm->flags = flags;
    mm->devs.assign(devs, devs + mm->ndev);
    mm->prog.assign(mm->ndev, nullptr);
    mm->slab.assign(nslabs, nullptr);
    mm->cuts.resize(nslabs + 1);
    mm->offs.assign(nslabs, 0);
    mm->up_dirty.assign(mm->ndev, 0);
    mm->h_stage.assign(mm->ndev, nullptr);
    mm->h_stage_cap.assign(mm->ndev, 0);
    mm->worker_err.assign(mm->ndev, std::string());
    slab_cuts(lat->n[2], nslabs, mm->cuts.data());
    int rc = 0;
    for (int w = 0; w < mm->ndev && !rc; w++) rc = gsdf_program_create_on(mm->devs[w], blob, blob_bytes, aux, aux_floats, &mm->prog[w]);
    for (int j = 0; j < nslabs && !rc; j++) {
        if (mm->cuts[j + 1] <= mm->cuts[j]) { rc = fail(GSDF_EINVAL, "internal: empty Z-slab %d", j); break; }
        rc = mesh_begin_prio(mm->prog[j % mm->ndev], lat, mm->cuts[j], mm->cuts[j + 1], flags, nullptr, multi_slab_priority(mm, j), &mm->slab[j]);
        if (!rc) mm->slab[j]->pdl_chain = multi_slab_pdl(mm, j);
    }
    if (rc) { gsdf_multi_destroy(mm); return rc; }
    // totals of the construction render
    mm->ntri = mm->evals = mm->pruned = 0;
    for (int j = 0; j < nslabs; j++) {
        mm->offs[j] = mm->ntri;
        mm->ntri += mm->slab[j]->ntri; mm->evals += mm->slab[j]->evals; mm->pruned += mm->slab[j]->pruned;
    }
    mm->rendered = true;
    for (int w = 1; w < mm->ndev; w++) mm->threads.emplace_back(multi_worker_main, mm, w);
    *out = mm;
    return 0;
}

int gsdf_multi_update(gsdf_multimesher *mm, const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats) {
    if (!mm) return fail(GSDF_EINVAL, "gsdf_multi_update: NULL handle");
    int rc = check_program_blob(blob, blob_bytes, aux, aux_floats, 3);
    if (rc) return rc;
    mm->up_blob.assign((const uint8_t *)blob, (const uint8_t *)blob + blob_bytes);
    mm->up_aux.assign(aux, aux + aux_floats);
    for (auto &d : mm->up_dirty) d = 1;
    return 0;
}

int gsdf_multi_specialize(gsdf_multimesher *mm) {
    if (!mm) return fail(GSDF_EINVAL, "gsdf_multi_specialize: NULL handle");
    for (int w = 0; w < mm->ndev; w++) {
        CU(use_device(mm->devs[w]));
        const int rc = program_specialize(mm->prog[w]);  // compiled once per structure: the other devices share the code
        if (rc) return rc;
    }
    return 0;
}

int64_t gsdf_multi_render(gsdf_multimesher *mm, float *tri9, size_t max_tris) {
    if (!mm) return fail(GSDF_EINVAL, "gsdf_multi_render: NULL handle");
    mm->dst = tri9; mm->max_tris = tri9 ? max_tris : 0;
    mm->dst_pinned = false;
    if (tri9) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, tri9) == cudaSuccess) mm->dst_pinned = at.type == cudaMemoryTypeHost;
        else (void)cudaGetLastError();
    }
    for (auto &c : mm->count) c.store(-1, std::memory_order_relaxed);
    mm->timeline.assign(2 + 2 * (size_t)mm->nslabs, 0.0);
    mm->t_call = std::chrono::steady_clock::now();
    mm->abort_rc.store(0);
    mm->job_done.store(0);
    if (mm->ndev > 1) {
        {
            std::lock_guard<std::mutex> lk(mm->mu);
            mm->job_seq.fetch_add(1, std::memory_order_release);
        }
        mm->cv.notify_all();
    }
    multi_worker_job(mm, 0);  // worker 0 is the calling thread
    while (mm->job_done.load(std::memory_order_acquire) < mm->ndev) std::this_thread::yield();
    if (const int rc = mm->abort_rc.load()) {
        for (const auto &e : mm->worker_err) if (!e.empty()) return fail(rc, "%s", e.c_str());
        return fail(rc, "multi-device render failed");
    }
    mm->ntri = mm->evals = mm->pruned = 0;
    float ms = 0;
    for (int j = 0; j < mm->nslabs; j++) {
        mm->offs[j] = mm->ntri;
        mm->ntri += mm->slab[j]->ntri; mm->evals += mm->slab[j]->evals; mm->pruned += mm->slab[j]->pruned;
    }
    for (int w = 0; w < mm->ndev; w++) {  // device time of worker w: first slab enqueued -> last slab finished
        cudaSetDevice(mm->devs[w]);
        for (int j = w; j < mm->nslabs; j += mm->ndev) {
            float t = 0;
            if (cudaEventElapsedTime(&t, mm->slab[w]->ev[0], mm->slab[j]->ev[4]) == cudaSuccess) ms = std::max(ms, t);
            else (void)cudaGetLastError();
        }
    }
    mm->device_ms = ms;
    static const bool dbg_stamps = getenv("GSDF_MULTI_DEBUG") != nullptr;  // profiles: the slabs' in-graph stage stamps on one clock
    if (dbg_stamps) {
        unsigned long long t0 = ~0ull;
        for (int j = 0; j < mm->nslabs; j++) for (int k = 0; k < kMeshStamps; k++) if (mm->slab[j]->h_stamp[k]) t0 = std::min(t0, mm->slab[j]->h_stamp[k]);
        for (int j = 0; j < mm->nslabs; j++) {
            const unsigned long long *t = mm->slab[j]->h_stamp;
            auto us = [&](int k) { return t[k] ? (double)(t[k] - t0) * 1e-3 : -1.0; };
            fprintf(stderr, "  slab %d stamps us: prune %.1f lists %.1f eval %.1f count %.1f scan %.1f emit %.1f end %.1f\n", j, us(0), us(1), us(2), us(3), us(4), us(5), us(7));
        }
    }
    mm->timeline.back() = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - mm->t_call).count();
    mm->rendered = true;
    mm->read_pos = 0;
    mm->delivered = tri9 != nullptr && mm->ntri <= max_tris;
    if (tri9 && mm->ntri > max_tris) return fail(GSDF_ESHORT, "destination holds %zu triangles, the render produced %llu", max_tris, (unsigned long long)mm->ntri);
    return (int64_t)mm->ntri;
}

int64_t gsdf_multi_read(gsdf_multimesher *mm, float *tri9, size_t max_tris) {
    if (!mm || !tri9) return fail(GSDF_EINVAL, "gsdf_multi_read: NULL argument");
    if (max_tris < 5) return fail(GSDF_ESHORT, "short buffer");  // flatrenderer.go:187
    uint64_t got = 0;
    while (got < max_tris && mm->read_pos < mm->ntri) {
        // the slab that holds triangle read_pos
        int j = mm->nslabs - 1;
        while (j > 0 && mm->offs[j] > mm->read_pos) j--;
        gsdf_mesher *m = mm->slab[j];
        const uint64_t in_slab = mm->read_pos - mm->offs[j];
        const uint64_t n = std::min<uint64_t>(m->ntri - in_slab, max_tris - got);
        if (n == 0) break;
        CU(use_device(m->device));
        CU(cudaMemcpy(tri9 + 9 * got, m->d_tris + 9 * in_slab, n * 9 * sizeof(float), cudaMemcpyDeviceToHost));
        got += n; mm->read_pos += n;
    }
    return (int64_t)got;  // 0 = io.EOF
}

int gsdf_multi_rebalance(gsdf_multimesher *mm, int rounds) {
    if (!mm) return fail(GSDF_EINVAL, "gsdf_multi_rebalance: NULL handle");
    int changed = 0;
    for (int r = 0; r < rounds; r++) {
        std::vector<double> cost(mm->nslabs);
        for (int j = 0; j < mm->nslabs; j++) cost[j] = (double)mm->slab[j]->evals;
        std::vector<int32_t> cuts(mm->nslabs + 1);
        // One device pipelining its read-back: the first slab's compute and the last slab's copy are the two parts of the
        // render nothing else runs under, so those two slabs get half a share of the work. Several devices: equal shares.
        std::vector<double> share(mm->nslabs, 1.0);
        if (mm->ndev == 1 && mm->nslabs >= 3) share.front() = share.back() = 0.5;
        slab_rebalance(mm->lat.n[2], mm->nslabs, mm->cuts.data(), cost.data(), cuts.data(), share.data());
        if (cuts == mm->cuts) break;
        for (int j = 0; j < mm->nslabs; j++) {
            if (cuts[j] == mm->cuts[j] && cuts[j + 1] == mm->cuts[j + 1]) continue;
            gsdf_mesher *fresh = nullptr;
            const int rc = mesh_begin_prio(mm->prog[j % mm->ndev], &mm->lat, cuts[j], cuts[j + 1], mm->flags, nullptr, multi_slab_priority(mm, j), &fresh);
            if (rc) return rc;  // the old partition stays valid up to slab j; the caller sees the error
            fresh->pdl_chain = multi_slab_pdl(mm, j);
            gsdf_mesh_destroy(mm->slab[j]);
            mm->slab[j] = fresh;
        }
        mm->cuts = cuts;
        changed = 1;
    }
    mm->ntri = mm->evals = mm->pruned = 0;
    for (int j = 0; j < mm->nslabs; j++) {
        mm->offs[j] = mm->ntri;
        mm->ntri += mm->slab[j]->ntri; mm->evals += mm->slab[j]->evals; mm->pruned += mm->slab[j]->pruned;
    }
    mm->read_pos = 0;
    return changed;
}

int gsdf_multi_timeline(const gsdf_multimesher *mm, double *us, int max_entries) {
    if (!mm || !us) return fail(GSDF_EINVAL, "gsdf_multi_timeline: NULL argument");
    const int n = (int)std::min<size_t>(mm->timeline.size(), (size_t)std::max(max_entries, 0));
    for (int i = 0; i < n; i++) us[i] = mm->timeline[i];
    return (int)mm->timeline.size();
}

int gsdf_multi_rewind(gsdf_multimesher *mm) {
    if (!mm) return fail(GSDF_EINVAL, "gsdf_multi_rewind: NULL handle");
    mm->read_pos = 0;
    return 0;
}

int gsdf_multi_stats(const gsdf_multimesher *mm, uint64_t *evals, uint64_t *pruned, uint64_t *tris, float *device_ms) {
    if (!mm) return fail(GSDF_EINVAL, "gsdf_multi_stats: NULL handle");
    if (evals) *evals = mm->evals;
    if (pruned) *pruned = mm->pruned;
    if (tris) *tris = mm->ntri;
    if (device_ms) *device_ms = mm->device_ms;
    return 0;
}

int gsdf_multi_slabs(const gsdf_multimesher *mm, int32_t *cuts, int32_t *devices, int max_slabs) {
    if (!mm) return fail(GSDF_EINVAL, "gsdf_multi_slabs: NULL handle");
    for (int j = 0; j < mm->nslabs && j < max_slabs; j++) {
        if (cuts) cuts[j] = mm->cuts[j];
        if (devices) devices[j] = mm->devs[j % mm->ndev];
    }
    if (cuts && mm->nslabs < max_slabs) cuts[mm->nslabs] = mm->cuts[mm->nslabs];
    return mm->nslabs;
}

void gsdf_multi_destroy(gsdf_multimesher *mm) {
    if (!mm) return;
    {
        std::lock_guard<std::mutex> lk(mm->mu);
        mm->quit.store(true);
    }
    mm->cv.notify_all();
    for (auto &t : mm->threads) t.join();
    for (auto *m : mm->slab) if (m) gsdf_mesh_destroy(m);
    for (auto *p : mm->prog) if (p) gsdf_program_destroy(p);
    for (auto *h : mm->h_stage) if (h) cudaFreeHost(h);
    if (mm->copyk_stream) cudaStreamDestroy(mm->copyk_stream);
    delete mm;
    (void)cudaGetLastError();
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
    k_scan_lookback<<<(unsigned)ntiles, kThreads, 0, st>>>(data, n, d->d_scanstate, d->d_ctr, d->scan_epoch, reinterpret_cast<unsigned long long *>(d->d_ctr + 2), nullptr);
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
    if ((rc = launch_dc(p, g, ((uint64_t)G.ncell + 3) / 4, st, nullptr))) return rc;
    k_dc_flags<<<grid_for(p->sms, G.ncell, 256), 256, 0, st>>>(d->d_dist, G.ncell, G.res, d->d_eidx);
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
    k_dc_compact<<<grid_for(p->sms, G.ncell, 256), 256, 0, st>>>(d->d_dist, d->d_eidx, G.ncell, G.res, d->d_cubekey);
    CU(cudaGetLastError());
    // RenderAll: origin + edge ends (dual_contour.go:85-107)
    g.mode = 1; g.cubekey = d->d_cubekey; g.ncubes = nc; g.dc4 = d->d_dc4;
    if ((rc = launch_dc(p, g, nc, st, nullptr))) return rc;
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
        if ((rc = launch_dc(p, g, (uint64_t)nc * 6, st, nullptr))) return rc;
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
    p->evals += d->evals;
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
    int rc = ensure_device(p->device);
    if (rc) return rc;
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
    CU(use_device(d->device));
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
static int64_t stl_from_device(int sms, const float *d_tri9, uint64_t n, uint8_t *&d_stl, size_t &stl_cap, void *dst, size_t dst_bytes, cudaStream_t st) {
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
    k_stl_pack<<<grid_for(sms, n, kThreads), kThreads, 0, st>>>(d_tri9, n, d_stl + 96);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(dst, d_stl + 12, bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return (int64_t)bytes;
}

int64_t gsdf_mesh_stl(gsdf_mesher *m, void *dst, size_t dst_bytes) {
    if (!m) return fail(GSDF_EINVAL, "NULL mesher");
    if (m->pending) return fail(GSDF_EINVAL, "a render is in flight on this mesher: call gsdf_mesh_rerun_end first");
    CU(use_device(m->device));
    DevInfo di;
    int rc = device_info(m->device, &di);
    if (rc) return rc;
    return stl_from_device(di.sms, m->d_tris, m->ntri, m->d_stl, m->stl_cap, dst, dst_bytes, m->stream);
}

int64_t gsdf_stl_pack(const float *tri9, size_t n, void *dst, size_t dst_bytes) {
    if (n == 0) return fail(GSDF_EEMPTY, "empty triangle slice");
    if (!tri9) return fail(GSDF_EINVAL, "NULL triangles");
    int rc = ensure_device(default_device());
    if (rc) return rc;
    DevInfo di;
    if ((rc = device_info(default_device(), &di))) return rc;
    float *d_t = nullptr;
    uint8_t *d_stl = nullptr;
    size_t cap = 0;
    cudaError_t e = cudaMalloc((void **)&d_t, n * 36);
    if (e != cudaSuccess) return fail(GSDF_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
    e = cudaMemcpy(d_t, tri9, n * 36, cudaMemcpyHostToDevice);
    int64_t r = e == cudaSuccess ? stl_from_device(di.sms, d_t, n, d_stl, cap, dst, dst_bytes, 0) : fail(GSDF_ECUDA, "H2D: %s", cudaGetErrorString(e));
    cudaFree(d_t);
    cudaFree(d_stl);
    return r;
}

}  // extern "C"

extern "C" {

// WriteBinarySTL of a multi-device render: every slab packs its own records on its device; the host buffer receives the
// header once and the slabs' records at 84 + 50 * (triangles before the slab).
int64_t gsdf_multi_stl(gsdf_multimesher *mm, void *dst, size_t dst_bytes) {
    if (!mm) return fail(GSDF_EINVAL, "gsdf_multi_stl: NULL handle");
    const uint64_t n = mm->ntri;
    if (n == 0) return fail(GSDF_EEMPTY, "empty triangle slice");                                               // stl.go:16-18
    if (n > 0xffffffffull) return fail(GSDF_EINVAL, "amount of triangles in model exceeds STL design limits");  // stl.go:21-23
    const size_t bytes = 84 + 50 * (size_t)n;
    if (!dst || dst_bytes < bytes) return fail(GSDF_ELEN, "STL buffer needs %zu bytes", bytes);
    uint8_t *out = static_cast<uint8_t *>(dst);
    std::memset(out, 0, 84);
    const uint32_t cnt = (uint32_t)n;
    std::memcpy(out + 80, &cnt, 4);
    for (int j = 0; j < mm->nslabs; j++) {  // enqueue every slab's packing and copy, then wait: devices work concurrently
        gsdf_mesher *m = mm->slab[j];
        if (m->ntri == 0) continue;
        CU(use_device(m->device));
        DevInfo di;
        int rc = device_info(m->device, &di);
        if (rc) return rc;
        if ((rc = grow(m->d_stl, m->stl_cap, 50 * (size_t)m->ntri + 16))) return rc;
        k_stl_pack<<<grid_for(di.sms, m->ntri, kThreads), kThreads, 0, m->stream>>>(m->d_tris, m->ntri, m->d_stl);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out + 84 + 50 * (size_t)mm->offs[j], m->d_stl, 50 * (size_t)m->ntri, cudaMemcpyDeviceToHost, m->stream));
    }
    for (int j = 0; j < mm->nslabs; j++) {
        CU(use_device(mm->slab[j]->device));
        CU(cudaStreamSynchronize(mm->slab[j]->stream));
    }
    return (int64_t)bytes;
}

}  // extern "C"
