/*
 * gsdf_b200.h -- C ABI of libgsdfb200.so: the B200 (sm_100a) SDF evaluator + mesher behind gsdf's own API.
 *
 * This is the drop-in boundary for the hot path
 *     glbuild.Shader3D tree -> gleval.SDF3.Evaluate -> glrender.Renderer.ReadTriangles -> glrender.WriteBinarySTL
 * Every entry point names the reference interface it replaces (paths relative to the soypat/gsdf repository).
 * A Go maintainer binds these with cgo (INTEGRATION.md shows the stub); tests bind them with ctypes.
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every call returns 0 / a count on success, a negative
 *     gsdf_status on failure; gsdf_last_error() returns a thread-local message for the last failure.
 *   - the library never retains caller pointers after a call returns (same contract as
 *     gleval/gpu_cgo.go:194-258, which copies in and out on every Evaluate).
 *   - handles are single-caller (like gleval.SDF3Compute, gleval/gpu.go:82-103); distinct handles may be
 *     used from distinct threads.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with GSDF_ECUDA.
 */
#ifndef GSDF_B200_H
#define GSDF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GSDF_OK = 0,
    GSDF_EINVAL = -1,   /* bad argument */
    GSDF_ELEN = -2,     /* gleval.errMismatchBufferLength (gleval/gleval.go:48) */
    GSDF_EEMPTY = -3,   /* gleval.errEmptyBuffers (gleval/gleval.go:47); empty model in WriteBinarySTL (stl.go:16) */
    GSDF_ECUDA = -4,    /* CUDA runtime error / no device */
    GSDF_ENOMEM = -5,
    GSDF_EPROGRAM = -6, /* malformed node program */
    GSDF_ESHORT = -7,   /* io.ErrShortBuffer: triangle buffer < 5 (octreerenderer.go:132, flatrenderer.go:187) */
    GSDF_ERES = -8,     /* "resolution not fine enough for marching cubes" (flatrenderer.go:53, octreerenderer.go:232) */
    GSDF_EUNSUPPORTED = -9 /* an optional facility is not available here (run-time compilation); the caller carries on without it */
} gsdf_status;

typedef struct gsdf_program gsdf_program; /* a compiled tree resident on one device */
typedef struct gsdf_mesher gsdf_mesher;   /* one Renderer instance (glrender.Renderer) */

/* Library / device ----------------------------------------------------------------------------------- */
const char *gsdf_version(void);
const char *gsdf_last_error(void);
/* Number of CUDA devices visible, or <0. */
int gsdf_device_count(void);
/* Select the device that handles created LATER BY THE CALLING THREAD live on (one process per GPU: pass LOCAL_RANK).
 * Replaces gleval.Init1x1GLFW (gleval/gpu.go:21) as the "bring the GPU up" call. The default is per thread; callers whose
 * threads are not theirs to pin (goroutines) use gsdf_program_create_on / gsdf_multi_begin, which take the device
 * explicitly. Every handle remembers its device: calls on a handle may come from any thread. */
int gsdf_set_device(int device);
/* Pinned (page-locked) host memory. Host pointers passed to gsdf_eval3/2, gsdf_mesh_read*, gsdf_multi_render may be any
 * memory; when they point into memory from gsdf_host_alloc (or memory the caller registered with cudaHostRegister) the
 * transfers run by DMA straight from / into the caller's buffer instead of through the library's staging buffers. A Go
 * caller wraps the pointer with unsafe.Slice. */
void *gsdf_host_alloc(size_t bytes);
void gsdf_host_free(void *p);

/* Program -------------------------------------------------------------------------------------------- */
/* Upload a flattened tree (include/gsdf_program.h). Replaces Programmer.WriteComputeSDF3 + NewComputeGPUSDF3
 * (glbuild/glbuild.go:175, gleval/gpu.go:35): "compile" is an upload of a few KB, done once.
 * blob = gsdf_program_header followed by header.nchunks 16-byte chunks; aux = side buffer of floats.
 * The blob is validated completely before any device work (GSDF_EPROGRAM otherwise): opcodes and lengths, aux ranges,
 * guard kinds and targets, the distance / position stack discipline against the header's slot counts, and that every
 * region a guard can skip leaves exactly the value its combiner consumes. */
int gsdf_program_create(const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats, gsdf_program **out);
/* Same on an explicitly named device (no per-thread state involved). */
int gsdf_program_create_on(int device, const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats, gsdf_program **out);
/* Re-upload a (re-)flattened tree of the same dimension into an existing handle: device buffers, stream and scheduler
 * are reused, so an edited tree costs one small host->device copy (the GL path recompiles its shader instead,
 * gleval/gpu.go:35-54). Renderers bound to the handle see the new tree on their next run. */
int gsdf_program_update(gsdf_program *p, const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats);
/* Compiles kernels specialised for this program's instruction stream (NVRTC, loaded at run time) and uses them for the
 * lattice evaluation and the prune-centre passes of every renderer built on the program: what constructing the GPU evaluator
 * does in the reference (it generates and compiles a GLSL compute shader per tree, glbuild/glbuild.go:175-214,
 * gleval/gpu.go:35-54). The program becomes straight-line code around the interpreter's own opcode bodies, so results stay
 * bit-identical; only the structure is compiled in (opcodes, flags, jump targets), operands are still read from the program, so
 * gsdf_program_update with new parameters of the same tree keeps the specialisation and an update that changes the structure
 * drops it (call again). Compiled code is cached per structure for the life of the process (one compilation of ~2 s per
 * distinct tree shape). Returns 0, or GSDF_EUNSUPPORTED when NVRTC is not available / GSDF_JIT=0 / the program is 2-D: the
 * interpreter kernels then keep running -- same results, slower. */
int gsdf_program_specialize(gsdf_program *p);
/* 1 when the specialised kernels are in use for the program's current structure. */
int gsdf_program_is_specialized(const gsdf_program *p);
/* The compilation step alone, without a device (build checks, tests): CUBIN size in bytes or a negative gsdf_status. */
int64_t gsdf_jit_compile(const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats);
/* What gsdf_program_create / _update put on the device for a flattened program, without a device (tests, debugging): the
 * instruction chunks (16 bytes each, header stripped) followed by the side buffer, to which the library appends operand
 * tables it derives itself -- today Sincos(angle * i), i = 0..ncirc, for every circarray (gsdf.go circarray.Evaluate,
 * cpu_evaluators.go:1056-1078), whose position goes into the unused fourth operand word of GSDF_OP_CIRC_ENTER. Returns the
 * image size in bytes (the side buffer starts at 16 * nchunks); copies it when image != NULL and image_bytes suffices,
 * GSDF_ESHORT when it does not. */
int64_t gsdf_program_device_image(const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats, void *image, size_t image_bytes);
void gsdf_program_destroy(gsdf_program *p);
/* gleval Evaluations() counter (gleval/cpu.go:126, gleval/gpu.go:80): points evaluated through this handle -- host and
 * device Evaluate calls, lattice and image evaluations, and the evaluations of the renderers bound to it (the reference's
 * renderers call sdf.Evaluate, so its counter includes them too). */
uint64_t gsdf_program_evaluations(const gsdf_program *p);

/* Evaluate ------------------------------------------------------------------------------------------- */
/* gleval.SDF3.Evaluate(pos []ms3.Vec, dist []float32, userData) (gleval/gleval.go:15-24). HOST pointers:
 * pos_xyz is n AoS float3 (ms3.Vec, 12 B), dist is n float32. The call is pipelined: the batch is cut into chunks that
 * move host -> device, through the kernel and device -> host on three rotating streams, so the copy of chunk i+1 and the
 * read-back of chunk i-1 run under the kernel of chunk i (the GL path, gleval/gpu_cgo.go:194-258, uploads, dispatches and
 * reads back serially). Pinned caller memory (gsdf_host_alloc) is transferred in place, other memory through the
 * handle's pinned staging. n==0 -> GSDF_EEMPTY (gleval/cpu.go:97-99). The Go shim checks len(pos)!=len(dist) ->
 * GSDF_ELEN before calling. */
int gsdf_eval3(gsdf_program *p, const float *pos_xyz, float *dist, size_t n);
/* gleval.SDF2.Evaluate (gleval/gleval.go:28-37): pos_xy is n AoS float2 (ms2.Vec, 8 B). */
int gsdf_eval2(gsdf_program *p, const float *pos_xy, float *dist, size_t n);
/* Same kernels on DEVICE pointers already resident in HBM (no copies); stream is a cudaStream_t or NULL. */
int gsdf_eval3_device(gsdf_program *p, const float *d_pos_xyz, float *d_dist, size_t n, void *stream);
int gsdf_eval2_device(gsdf_program *p, const float *d_pos_xy, float *d_dist, size_t n, void *stream);

/* Dense lattice ---------------------------------------------------------------------------------------- */
typedef struct {
    float origin[3]; /* lattice corner (0,0,0): (Bounds() scaled 1.01 about its centre).Min */
    float res;       /* cube resolution */
    int32_t n[3];    /* cells per axis; corners are (n+1) per axis */
} gsdf_lattice;

/* FlatRenderer.Reset's lattice (glrender/flatrenderer.go:47-56): bb*1.01, n=ceil(size/res). GSDF_ERES if n<=0. */
int gsdf_lattice_from_bounds(const float bbmin[3], const float bbmax[3], float res, gsdf_lattice *out);
/* makeICube's level count (glrender/octreerenderer.go:222-235): ceil(log2(longAxis/res))+1 on the 1.01-scaled box;
 * GSDF_ERES when it is <= 1 ("resolution not fine enough for marching cubes"), GSDF_EINVAL for res <= 0 / NaN / Inf. */
int gsdf_octree_levels(const float bbmin[3], const float bbmax[3], float res);
/* FlatRenderer.evalGrid / evalKRange (glrender/flatrenderer.go:103-182) for corner planes k in [k0,k1):
 * positions origin + float32(i)*res are synthesised in-kernel; (n0+1)*(n1+1)*(k1-k0) distances, x fastest, are
 * written to `dist` (HOST pointer) or, if dist is NULL, only computed (timing). */
int gsdf_grid_eval(gsdf_program *p, const gsdf_lattice *lat, int k0, int k1, float *dist);
/* Same, into a DEVICE buffer with row pitch (n0+1) floats. */
int gsdf_grid_eval_device(gsdf_program *p, const gsdf_lattice *lat, int k0, int k1, float *d_dist, void *stream);

/* Mesher --------------------------------------------------------------------------------------------------- */
enum {
    GSDF_MESH_PRUNE = 1u << 0,      /* octree level-3 prune (octreerenderer.go:180-191,240-284); off = FlatRenderer */
    GSDF_MESH_KEEP_CASES = 1u << 1, /* also keep the 8-bit cube-case index per cell (parity checks) */
    GSDF_MESH_KEEP_GRID = 1u << 2,  /* keep the distance lattice readable through gsdf_mesh_grid */
    GSDF_MESH_STAGE_TIMING = 1u << 3, /* time every stage (gsdf_mesh_timings [0..3]): launches eagerly with events between the
                                        stages; without it steady-state reruns replay one CUDA graph and only [4] is filled */
    GSDF_MESH_PRUNE_LITERAL = 1u << 4 /* with GSDF_MESH_PRUNE: margin 1 at every level-3 cube (see gsdf_prune_plan) */
};

/* The coarse-to-fine prune. The reference's rule (octreePrunea, octreerenderer.go:240-284) drops a cube when
 * |d(centre)| >= size * sqrt3/2. Its scheduler (octreerenderer.go:94-104,136-151) applies the rule first to the cubes of
 * level top-4 (4096 cubes fill its 4680-cube prune buffer) and to level-3 cubes only for the part of the model that is
 * still unrendered when that buffer drains -- which part that is depends on buffer sizes and un-vendored helpers. On
 * fields that are not 1-Lipschitz (smooth blends, knurls) the literal rule at level 3 everywhere drops a few cells that
 * hold surface, which the reference's own runs do not lose (README.md:152: 309,872 triangles from both renderers).
 * A plan is a list of levels, coarse to fine (level L = cubes of 2^(L-1) cells, aligned to the lattice origin like
 * ms3.Octree cubes); a cube is kept iff |d(centre)| < margin * size * sqrt3/2 and only the children of kept cubes are
 * looked at on the next level. A plan ends with level 3 (the marching-cubes stage works on 4-cell blocks), optionally
 * followed by level 2: the 2-cell cubes inside kept blocks decide which lattice corners are evaluated at all and which
 * cells the marching-cubes stage looks at (levels 3 and 2 run as one launch).
 * Default plan (GSDF_MESH_PRUNE): level 3 with margin GSDF_PRUNE_MARGIN_DEFAULT, preceded by a coarse level on large
 * lattices; it reproduces the dense sweep on every scene of the reference's examples, the README's two known answers
 * included (level 2 is left to explicit plans: with margin 1.25 it loses triangles on the steepest of those fields).
 * GSDF_MESH_PRUNE_LITERAL: the same levels with margin 1 (bit-equal to the dense sweep on 1-Lipschitz fields). */
#define GSDF_PRUNE_MARGIN_DEFAULT 1.25f
#define GSDF_PRUNE_MAX_LEVELS 4
typedef struct {
    int32_t nlevels;
    int32_t level[GSDF_PRUNE_MAX_LEVELS];  /* strictly descending, each in [2, 12], ending with 3 or with 3, 2 */
    float margin[GSDF_PRUNE_MAX_LEVELS];   /* >= 1 */
} gsdf_prune_plan;
/* The plan gsdf_mesh_begin uses for `flags` on this lattice. */
int gsdf_prune_plan_default(const gsdf_lattice *lat, unsigned flags, gsdf_prune_plan *out);
/* glrender.NewOctreeRenderer / FlatRenderer.Reset (octreerenderer.go:45, flatrenderer.go:37) on cells
 * cz in [cz0,cz1) of the lattice (Z-slab; pass 0,n[2] for everything). The whole slab is meshed on the device
 * inside this call; triangles stay in HBM until read. */
int gsdf_mesh_begin(gsdf_program *p, const gsdf_lattice *lat, int cz0, int cz1, unsigned flags, gsdf_mesher **out);
/* Same with an explicit prune plan (implies GSDF_MESH_PRUNE). */
int gsdf_mesh_begin_plan(gsdf_program *p, const gsdf_lattice *lat, int cz0, int cz1, unsigned flags, const gsdf_prune_plan *plan,
                         gsdf_mesher **out);
/* Re-run the same slab on the same handle, reusing every device buffer (Renderer.Reset semantics). */
int gsdf_mesh_rerun(gsdf_mesher *m);
/* Point the mesher at another (3D) program on the same device, keeping lattice and buffers: Renderer.Reset with a
 * new SDF (octreerenderer.go:72, flatrenderer.go:37). The caller keeps ownership of both programs. */
int gsdf_mesh_set_program(gsdf_mesher *m, gsdf_program *p);
/* Renderer.ReadTriangles(dst []ms3.Triangle) (glrender/glrender.go:11-13): copies up to max_tris triangles
 * (9 floats each, vertex order as marchcubes.go:64-68) in FlatRenderer order (cell index x fastest,
 * flatrenderer.go:208-212). Returns the count (0 = io.EOF), GSDF_ESHORT if max_tris < 5. */
int64_t gsdf_mesh_read(gsdf_mesher *m, float *tri9, size_t max_tris);
/* Asynchronous form: enqueues the copy of up to max_tris triangles (from the current read position) on the mesher's
 * copy stream and returns their count at once; tri9 must stay valid (and should be pinned) until gsdf_mesh_wait.
 * Lets the device->host transfer of one slab overlap the kernels of the next one (separate meshers per slab). */
int64_t gsdf_mesh_read_async(gsdf_mesher *m, float *tri9, size_t max_tris);
int gsdf_mesh_wait(gsdf_mesher *m);
/* Split form of gsdf_mesh_rerun for pipelines of several meshers (Z-slabs of one lattice, or successive renders):
 * _begin enqueues the render and returns at once; _end waits for it, reads its counters and finishes the bookkeeping
 * (re-emitting if the triangle buffer was too small). Between the two only gsdf_mesh_read_prefix_async and gsdf_mesh_wait
 * may be called on the handle. */
int gsdf_mesh_rerun_begin(gsdf_mesher *m);
int gsdf_mesh_rerun_end(gsdf_mesher *m);
/* Enqueues, behind the render enqueued last, the copy of the FIRST ntris triangles of the slab's buffer to tri9 (pinned
 * host memory) without knowing how many the render will produce -- the caller speculates (normally: the count of the
 * previous render) and checks gsdf_mesh_stats after gsdf_mesh_rerun_end; on a mismatch it re-reads with gsdf_mesh_read.
 * This takes the "how many triangles?" round trip out of the critical path: the copy of slab i runs under the kernels of
 * slab i+1 with no host synchronisation in between. Returns the number of triangles whose copy was enqueued (ntris clamped
 * to the buffer's capacity); gsdf_mesh_wait completes it. */
int64_t gsdf_mesh_read_prefix_async(gsdf_mesher *m, float *tri9, size_t ntris);
/* Device pointer to the slab's triangle buffer (9 floats per triangle) and its count, for on-device consumers. */
int gsdf_mesh_device_triangles(gsdf_mesher *m, const float **d_tri9, uint64_t *ntri);
/* Evaluations() / Octree.TotalPruned() / len(triangles) (gsdfaux/gsdfaux.go:219-226). evals counts the evaluations the render
 * executed: prune-cube centres + the listed lattice corners (whole quads of 4 corners, or -- specialised kernels on lattices
 * that list 2^19 quads or more -- half-quads of 2, which leaves out 5-9 % of the corners nobody reads). */
int gsdf_mesh_stats(const gsdf_mesher *m, uint64_t *evals, uint64_t *pruned_unit_cubes, uint64_t *tris);
/* Parity hooks: per-cell case indices (nx*ny*(cz1-cz0) bytes) and the corner lattice of the slab. */
int gsdf_mesh_cases(gsdf_mesher *m, uint8_t *cases, size_t nbytes);
int gsdf_mesh_grid(gsdf_mesher *m, float *grid, size_t nfloats);
/* Milliseconds of device time (CUDA events) the last begin/rerun spent in: [0] prune, [1] fine eval,
 * [2] classify+scan, [3] emit, [4] total. */
int gsdf_mesh_timings(const gsdf_mesher *m, float ms[5]);
void gsdf_mesh_destroy(gsdf_mesher *m);

/* Multi-device mesher ---------------------------------------------------------------------------------------------- */
/* One lattice meshed by several GPUs from ONE process: the device analogue of FlatRenderer.evalGrid's split of the corner
 * planes over goroutines (glrender/flatrenderer.go:103-141). The cell layers are cut into ndev * slabs_per_device Z-slabs
 * (one shared corner plane between neighbours, cuts on 4-layer block boundaries where the lattice allows), dealt round-robin
 * to the devices -- slab j lives on devs[j % ndev] -- so that slabs finish in roughly their output order and a lattice
 * whose surface is concentrated in a few layers still spreads over all devices. Every slab is an ordinary mesher with its
 * own stream; every device has its own copy of the program, one host worker thread, and pinned staging. There is no
 * collective: per-slab triangle buffers are concatenated in slab order in the caller's host buffer, which reproduces the
 * single-device output bit for bit. ndev == 1 with several slabs pipelines one device: the read-back of slab i runs under
 * the kernels of slab i+1 (the counts are read, not predicted). */
typedef struct gsdf_multimesher gsdf_multimesher;
int gsdf_multi_begin(int ndev, const int *devs, int slabs_per_device, const void *blob, size_t blob_bytes, const float *aux,
                     size_t aux_floats, const gsdf_lattice *lat, unsigned flags, gsdf_multimesher **out);
/* gsdf_program_update on every device's copy of the program. */
int gsdf_multi_update(gsdf_multimesher *mm, const void *blob, size_t blob_bytes, const float *aux, size_t aux_floats);
/* gsdf_program_specialize for the program of every device of the handle (one compilation, shared). 0 or GSDF_EUNSUPPORTED. */
int gsdf_multi_specialize(gsdf_multimesher *mm);
/* Render the lattice and deliver every triangle, in FlatRenderer order, to tri9 (HOST memory for max_tris triangles; pinned
 * memory is written by DMA directly). Returns the triangle count; if the buffer is too small nothing is copied and the
 * call fails with GSDF_ESHORT (gsdf_multi_stats then tells the count). tri9 == NULL renders without read-back. */
int64_t gsdf_multi_render(gsdf_multimesher *mm, float *tri9, size_t max_tris);
/* Renderer.ReadTriangles on the last render: up to max_tris triangles from the read position (0 = io.EOF). */
int64_t gsdf_multi_read(gsdf_multimesher *mm, float *tri9, size_t max_tris);
/* Host-clock timeline of the last gsdf_multi_render, microseconds from the call: us[0] = every slab of worker 0 enqueued,
 * us[1+2j] = slab j's counters seen, us[2+2j] = slab j's read-back enqueued (0 without a pinned destination), last entry =
 * everything delivered. Returns the number of entries (2 + 2 * slabs). For profiles, not for control flow. */
int gsdf_multi_timeline(const gsdf_multimesher *mm, double *us, int max_entries);
/* Moves the read position of gsdf_multi_read back to the first triangle of the last render. */
int gsdf_multi_rewind(gsdf_multimesher *mm);
/* Totals of the last render; device_ms = the longest device time over the slabs' streams (may be NULL). */
int gsdf_multi_stats(const gsdf_multimesher *mm, uint64_t *evals, uint64_t *pruned_unit_cubes, uint64_t *tris, float *device_ms);
/* The partition: nslabs+1 cell-layer cuts and the device of each slab; returns nslabs (arrays may be NULL). */
int gsdf_multi_slabs(const gsdf_multimesher *mm, int32_t *cuts, int32_t *devices, int max_slabs);
/* WriteBinarySTL of the last render packed on the devices (per-slab records concatenated behind one header). */
int64_t gsdf_multi_stl(gsdf_multimesher *mm, void *dst, size_t dst_bytes);
void gsdf_multi_destroy(gsdf_multimesher *mm);
/* The Z-slab cuts gsdf_multi_begin uses (host arithmetic): nslabs+1 ascending cell layers from 0 to nz. */
int gsdf_slab_cuts(int nz, int nslabs, int32_t *cuts);
/* Cuts that equalise a per-slab cost (host arithmetic): cost[j] is what slab j = [cuts[j], cuts[j+1]) cost in the last render
 * (evaluations executed, or milliseconds); it is spread evenly over the slab's layers and out[] is cut where the cumulative
 * cost reaches j/nslabs of the total. Pruned lattices put most of their work where the surface is, not where the volume is
 * (SURVEY 8e): equal layers are not equal work. */
int gsdf_slab_rebalance(int nz, int nslabs, const int32_t *cuts, const double *cost, int32_t *out);
/* Up to `rounds` rounds of gsdf_slab_rebalance on the evaluations each slab executed in its last render; slabs whose range
 * changes are rebuilt (and rendered once). Returns 1 if the partition changed, 0 if it was already balanced, <0 on error. */
int gsdf_multi_rebalance(gsdf_multimesher *mm, int rounds);

/* Dual contouring --------------------------------------------------------------------------------------------- */
/* glrender.DualContourRenderer (glrender/dual_contour.go:12-218) with its vertex placement strategies
 * (glrender/dual_contour_vertexplacement.go) and gleval.NormalsCentralDiff (gleval/gleval.go:53-108) on the device. */
typedef struct gsdf_dualcontour gsdf_dualcontour;
typedef enum {
    GSDF_DC_NAIVE = 0,                 /* DualContourNaive: mean of the edge crossings (dual_contour_test.go:355-389) */
    GSDF_DC_LEAST_SQUARES = 1,         /* DualContourLeastSquares{} (dual_contour_vertexplacement.go:16-138) */
    GSDF_DC_LEAST_SQUARES_CHISELED = 2 /* DualContourLeastSquares{Chiseled: true} (:21, :43-46, :118-121) */
} gsdf_dc_placer;
/* makeICube on Bounds().Add(-res/2) (dual_contour.go:31-36, octreerenderer.go:222-235): octree level count, GSDF_ERES if
 * <= 1; origin (optional) receives the octree origin. */
int gsdf_dc_levels(const float bbmin[3], const float bbmax[3], float res, float origin[3]);
/* DualContourRenderer.Reset(sdf, res, placer) + RenderAll: the whole mesh is built on the device inside this call.
 * bbmin/bbmax = sdf.Bounds(). Limits: at most 1024 cubes per axis (11 levels). */
int gsdf_dc_begin(gsdf_program *p, const float bbmin[3], const float bbmax[3], float res, int placer, gsdf_dualcontour **out);
/* Multi-GPU form: this handle owns part `part` of `nparts` (1, 2, 4 or 8) equal, contiguous ranges of the octree's BFS
 * cube order -- runs of top-level octants. It evaluates its octants plus a one-cube border on each side (the neighbour data the QEF
 * and the quads of its own cubes need), and emits only the quads of its own cubes; the parts' triangle buffers
 * concatenated in part order are bit-identical to the single-handle mesh. No collective on the data path. */
int gsdf_dc_begin_part(gsdf_program *p, const float bbmin[3], const float bbmax[3], float res, int placer, int part, int nparts,
                       gsdf_dualcontour **out);
/* The partition itself (host arithmetic, no device needed): keys = the part's range [keys[0], keys[1]) of the octree's
 * BFS cube order (index = per-level child indices of i3.Cube.Octree(), most significant level first); box = {lo.xyz,
 * hi.xyz}, the half-open range of cube indices whose origins the part evaluates (its octants plus a one-cube border on each side). */
int gsdf_dc_part_region(int levels, int part, int nparts, uint32_t keys[2], int32_t box[6]);
/* Re-run with the same parameters, reusing every device buffer. */
int gsdf_dc_rerun(gsdf_dualcontour *d);
/* RenderAll's result: copies up to max_tris triangles (9 floats each, cube order, two per quad) from the start of the
 * mesh; returns the number copied. */
int64_t gsdf_dc_read(gsdf_dualcontour *d, float *tri9, size_t max_tris);
int gsdf_dc_device_triangles(gsdf_dualcontour *d, const float **d_tri9, uint64_t *ntri);
/* stats = {octree levels, cubes kept by the prune (len(cubebuf)), cubes that received a vertex (len(Neighbors) > 0),
 * triangles, SDF evaluations the reference performs for this render, microseconds of device time}. */
int gsdf_dc_stats(const gsdf_dualcontour *d, uint64_t stats[6]);
void gsdf_dc_destroy(gsdf_dualcontour *d);

/* STL ---------------------------------------------------------------------------------------------------- */
/* glrender.WriteBinarySTL (glrender/stl.go:15-62): 80 zero bytes + u32 count + 50 B per triangle (unit normal,
 * 3 vertices, u16 0). Packs n HOST triangles into dst (needs 84+50*n bytes) on the device and returns the byte
 * count, so the caller does ONE write. n==0 -> GSDF_EEMPTY. */
int64_t gsdf_stl_pack(const float *tri9, size_t n, void *dst, size_t dst_bytes);
/* Same, straight from a mesher's device triangles (no host round trip of the 36 B/triangle form). */
int64_t gsdf_mesh_stl(gsdf_mesher *m, void *dst, size_t dst_bytes);

/* 2D image ------------------------------------------------------------------------------------------------ */
/* ImageRendererSDF2.Render's evaluation (glrender/image.go:76-105): dist[j*w+i] = sdf(x_i, y_j) with
 * x_i = float32(i)*dx + (min.x+dx/2), y_j = max.y - float32(j)*dy. dist is a HOST pointer (w*h floats). */
int gsdf_image_eval2(gsdf_program *p, const float bbmin[2], const float bbmax[2], int w, int h, float *dist);

/* Colour conversion of ImageRendererSDF2 (glrender/image.go:20-24,50-62) fused into the evaluation: the reference calls a
 * Go closure per pixel through img.Set (image.go:112-116); here the conversions gsdfaux offers are data, applied by
 * the kernel that produced the distance, and the image leaves the device as RGBA8 (4 B/pixel). */
typedef enum {
    GSDF_CONV_DEFAULT = 0,      /* NewImageRendererSDF2(nil): NaN/Inf red, d>0 white, else black (image.go:50-61) */
    GSDF_CONV_BW_LINEAR = 1,    /* blackAndWhiteLinearSmooth(edge) (gsdfaux/color.go:77-95); edge 0 = hard step (:97-102) */
    GSDF_CONV_INIGO_QUILEZ = 2, /* ColorConversionInigoQuilez(characteristicDistance) (gsdfaux/color.go:21-47) */
    GSDF_CONV_HSV_GRADIENT = 3  /* ColorConversionLinearGradient(len, c0, c1), general case (gsdfaux/color.go:51-73) */
} gsdf_conv_kind;
typedef struct {
    int32_t kind;
    float p[7];      /* BW_LINEAR: p0 = edge; INIGO_QUILEZ: p0 = 1/characteristicDistance; HSV: h0,s0,v0,h1,s1,v1,len */
    uint32_t c0, c1; /* HSV: end colours as little-endian RGBA8 bytes (R | G<<8 | B<<16 | A<<24) */
} gsdf_colorconv;
/* gsdfaux.ColorConversionInigoQuilez (color.go:21): inv = 1/characteristicDistance in float32. */
int gsdf_colorconv_inigo_quilez(float characteristic_distance, gsdf_colorconv *out);
/* gsdfaux.ColorConversionLinearGradient(gradientLength, c0, c1) (color.go:51): black -> white selects BW_LINEAR like the
 * reference (:52-54); other colours are converted to HSV on the host as colorToHSV does (:127-130). rgba = R|G<<8|B<<16|A<<24. */
int gsdf_colorconv_linear_gradient(float gradient_length, uint32_t rgba0, uint32_t rgba1, gsdf_colorconv *out);
/* ImageRendererSDF2.Render (glrender/image.go:76-118) with the conversion fused: rgba[j*w+i] = conv(sdf(x_i, y_j)),
 * positions as gsdf_image_eval2. conv NULL = GSDF_CONV_DEFAULT. rgba is a HOST pointer to w*h*4 bytes (image.RGBA.Pix). */
int gsdf_image_render2(gsdf_program *p, const float bbmin[2], const float bbmax[2], int w, int h, const gsdf_colorconv *conv,
                       uint8_t *rgba);
/* Same two calls into 16-byte aligned DEVICE buffers (w*h floats / w*h*4 bytes) on `stream` (cudaStream_t or NULL = the
 * handle's stream), no copies and no synchronisation: the image stays in HBM for an on-device consumer. */
int gsdf_image_eval2_device(gsdf_program *p, const float bbmin[2], const float bbmax[2], int w, int h, float *d_dist, void *stream);
int gsdf_image_render2_device(gsdf_program *p, const float bbmin[2], const float bbmax[2], int w, int h, const gsdf_colorconv *conv,
                              uint8_t *d_rgba, void *stream);

#ifdef __cplusplus
}
#endif
#endif
