/*
 * gsdf_oracle.h -- CPU ORACLE for the gsdf SDF-evaluate + mesh path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's CPU algorithm (soypat/gsdf, pure Go).  It exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can check and time the
 * CUDA path against it.  Nothing in the shipped product path (gsdf_b200/) may import, link or call it.
 *
 * PARITY PINNING STATUS
 *   pinned   : marching-cubes tables (checked against glrender/marchcubes.go:101-412 by tests/test_mc_tables.py
 *              when /root/reference is present), lattice formula (README.md:130 eval count 6,711,685+1),
 *              sphere KAT 41072 triangles (glrender/glrender_test.go:83-99), ISO thread sign KAT
 *              (forge/threads/threads_test.go:14-44), STL byte layout (glrender/stl.go:15-119).
 *   UNPINNED : last-bit behaviour of chewxy/math32 v1.11.1 (Sin/Cos/Tan/Atan/Atan2/Hypot) and of
 *              soypat/geometry v0.0.0-20251107203642-291c5648d529 helpers (ms3.Norm, Unit, Triangle.Normal,
 *              octree traversal order).  Those modules are not vendored under /root/reference and no Go
 *              toolchain exists in this image, so the reference itself cannot be run here.  The restatement
 *              follows their published algorithms (Go math / Cephes float32 ports; gonum r3-style vector ops).
 *
 * Each function cites the reference file:line it follows.
 */
#ifndef GSDF_ORACLE_H
#define GSDF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- tree table: one record per CSG node, raw struct fields exactly as the Go node types hold them ---- */
enum {
    /* 3D primitives (primitives.go) */
    GO_SPHERE = 1, GO_BOX, GO_CYLINDER, GO_HEX, GO_TORUS, GO_BOXFRAME,
    /* 3D operations (operations.go) */
    GO_UNION = 16, GO_DIFF, GO_INTERSECT, GO_XOR, GO_SMOOTH_UNION, GO_SMOOTH_DIFF, GO_SMOOTH_INTERSECT,
    GO_SCALE, GO_SYMMETRY, GO_TRANSFORM, GO_TRANSLATE, GO_OFFSET, GO_ARRAY, GO_ELONGATE, GO_SHELL,
    GO_CIRCARRAY, GO_TWIST, GO_BOUNDS3,
    /* 2D -> 3D (operations2d.go, forge/threads/threads.go) */
    GO_EXTRUDE = 40, GO_REVOLVE, GO_SCREW,
    /* 2D primitives (primitives2d.go) */
    GO_LINE2D = 64, GO_LINES2D, GO_ARC2D, GO_CIRCLE2D, GO_EQTRI2D, GO_RECT2D, GO_HEX2D, GO_OCT2D,
    GO_ELLIPSE2D, GO_POLY2D, GO_DIAMOND2D, GO_ROUNDX2D, GO_BEZIERQ2D,
    /* 2D operations (operations2d.go) */
    GO_UNION2D = 96, GO_DIFF2D, GO_INTERSECT2D, GO_XOR2D, GO_ARRAY2D, GO_OFFSET2D, GO_TRANSLATE2D,
    GO_ROTATE2D, GO_SYMMETRY2D, GO_ANNULUS2D, GO_CIRCARRAY2D, GO_SCALE2D, GO_TRANSLATEMULTI2D, GO_ELONGATE2D, GO_BOUNDS2
};

typedef struct {
    int32_t kind;       /* GO_* */
    int32_t nchild;     /* number of children */
    int32_t child_off;  /* index of first child id in the children[] array */
    int32_t aux_off;    /* float offset into aux[] (polygon vertices, line segments, displacements) */
    int32_t aux_cnt;    /* number of floats in aux */
    int32_t iparam[3];  /* integer fields (nx,ny,nz | nInst,circleDiv | mirror bits) */
    float   fparam[16]; /* float fields in the order the Go struct declares them */
} go_node;              /* 96 bytes */

typedef struct {
    const go_node *nodes;
    int32_t        nnodes;
    const int32_t *children;
    const float   *aux;
    int32_t        root;
} go_tree;

/* ---- float32 math restated from chewxy/math32 (exported so tests can check them against libm) ---- */
float go_sqrt(float x);
float go_hypot(float p, float q);
float go_atan(float x);
float go_atan2(float y, float x);
float go_sin(float x);
float go_cos(float x);
float go_tan(float x);
float go_acos(float x);
float go_cbrt(float x);
float go_log(float x);
float go_exp(float x);
float go_floor(float x);
float go_round(float x);
float go_min(float a, float b);
float go_max(float a, float b);

/* ---- evaluation: gleval.SDF3.Evaluate / SDF2.Evaluate on the CPU (cpu_evaluators.go) ---- */
/* returns 0 on success, <0 on malformed tree.  pos is AoS xyz (ms3.Vec) or xy (ms2.Vec). */
int go_eval3(const go_tree *t, const float *pos_xyz, float *dist, size_t n);
int go_eval2(const go_tree *t, const float *pos_xy, float *dist, size_t n);

/* ---- FlatRenderer (glrender/flatrenderer.go) ---- */
typedef struct {
    float origin[3];
    float res;
    int32_t n[3]; /* cells per axis; lattice is (n+1)^3 corners */
} go_lattice;

/* flatrenderer.go:47-56: bb scaled 1.01 about its centre, n = ceil(size/res). returns <0 if n<=0. */
int go_flat_lattice(const float bbmin[3], const float bbmax[3], float res, go_lattice *out);
/* octreerenderer.go:222-235: levels = ceil(log2(longAxis/res))+1 on the 1.01-scaled box. <0 on error. */
int go_octree_levels(const float bbmin[3], const float bbmax[3], float res);

/* flatrenderer.go:103-182: evaluate every lattice corner, nthreads k-slab workers, batches of batch points.
 * grid has (n0+1)(n1+1)(n2+1) floats, x fastest. returns evaluations performed or <0. */
int64_t go_flat_eval_grid(const go_tree *t, const go_lattice *lat, float *grid, int nthreads, int batch);

/* flatrenderer.go:186-256 + marchcubes.go:34-98: serial sweep over cells x-fastest.
 * tri9: up to max_tris*9 floats; cases: optional nx*ny*nz bytes (0 for rejected cells) or NULL.
 * blockmask: optional octree-prune mask over 4x4x4-cell blocks (1 = keep), NULL = FlatRenderer semantics.
 * returns number of triangles (counting continues past max_tris; only the first max_tris are stored). */
int64_t go_flat_march(const go_lattice *lat, const float *grid, float *tri9, int64_t max_tris, uint8_t *cases,
                      const uint8_t *blockmask);

/* The same sweep for cell layers [cz0,cz1) only (one rank's Z-slab); arrays still describe the whole lattice. */
int64_t go_flat_march_slab(const go_lattice *lat, const float *grid, float *tri9, int64_t max_tris, uint8_t *cases,
                           const uint8_t *blockmask, int cz0, int cz1);

/* octreerenderer.go:180-191,240-284: the level-3 prune rule on the flat lattice.
 * mask gets ceil(n/4)^3 bytes (x fastest): 1 if |d(centre)| < size*sqrt3/2 for the 4-cell cube, else 0.
 * returns the number of kept blocks or <0. */
int64_t go_flat_eval_planes(const go_tree *t, const go_lattice *lat, int k0, int k1, float *planes, int nthreads, int batch);
int64_t go_octree_prune_plan(const go_tree *t, const go_lattice *lat, int nlevels, const int *levels, const float *margins, uint8_t *mask,
                             int64_t *evals);
int64_t go_octree_prune_mask(const go_tree *t, const go_lattice *lat, uint8_t *mask);

/* marchcubes.go:34-73 single cube (exported for table/unit tests). p: 8 corners xyz, v: 8 values. returns ntri. */
int go_mc_cube(const float p[24], const float v[8], float tri9[45], int *case_index);

/* ---- dual contouring (gsdf_oracle_dc.c): glrender/dual_contour.go + dual_contour_vertexplacement.go ---- */
/* makeICube on Bounds().Add(-res/2) (dual_contour.go:31-34): level count (<0 = resolution too coarse) and octree origin. */
int go_dc_levels(const float bbmin[3], const float bbmax[3], float res, float origin[3]);
/* leastSquaresMGS64 (dual_contour_vertexplacement.go:148-223) on K <= 32 rows; A is K x 3 row-major */
int go_lsq_mgs64(int K, const float *A_rowmajor, const float *b, float x3[3]);
/* DualContourRenderer.Reset + RenderAll. placer: 0 DualContourNaive (dual_contour_test.go:355), 1 DualContourLeastSquares{},
 * 2 DualContourLeastSquares{Chiseled:true}. Returns the triangle count (only the first max_tris are stored) or <0.
 * stats (optional): {levels, cubes kept, cubes with neighbours, evaluations}. */
int64_t go_dual_contour(const go_tree *t, const float bbmin[3], const float bbmax[3], float res, int placer, float *tri9, int64_t max_tris,
                        int64_t *stats);

/* Same, also reporting per quad (= per two triangles) the BFS key of the cube that emitted it (multi-rank partition tests). */
int64_t go_dual_contour_ex(const go_tree *t, const float bbmin[3], const float bbmax[3], float res, int placer, float *tri9, int64_t max_tris,
                           int64_t *stats, uint32_t *quad_keys);

/* ---- STL (glrender/stl.go:15-62) ---- */
/* dst needs 84 + 50*ntri bytes. returns bytes written, or <0 (empty model is an error, stl.go:16). */
int64_t go_stl_write(const float *tri9, int64_t ntri, uint8_t *dst);
/* stl.go:175-225 minimal reader: returns ntri or <0; tri9 may be NULL to query the count. */
int64_t go_stl_read(const uint8_t *src, size_t nbytes, float *tri9, int64_t max_tris);

/* ---- ImageRendererSDF2 positions (glrender/image.go:76-105) ---- */
int go_image_eval2(const go_tree *t, const float bbmin[2], const float bbmax[2], int w, int h, float *dist);

/* Colour conversions of the 2-D image path: kind 0 = NewImageRendererSDF2(nil) (glrender/image.go:50-61), 1 =
 * blackAndWhiteLinearSmooth (gsdfaux/color.go:77-102), 2 = ColorConversionInigoQuilez (:21-47), 3 =
 * ColorConversionLinearGradient general case (:57-72). p / c0 / c1 as gsdf_colorconv (include/gsdf_b200.h). */
typedef struct { int32_t kind; float p[7]; uint32_t c0, c1; } go_colorconv;
uint32_t go_color_of(const go_colorconv *cc, float d);
void go_rgb_to_hsv(float r, float g, float b, float hsv[3]);     /* gsdfaux/color.go:192-217 */
/* ImageRendererSDF2.Render (glrender/image.go:76-118) into RGBA8 (image.RGBA.Pix order). */
int go_image_render2(const go_tree *t, const float bbmin[2], const float bbmax[2], int w, int h, const go_colorconv *cc, uint8_t *rgba);

/* tables, for cross-checking */
const int *go_mc_edge_table(void);     /* 256 */
const int8_t *go_mc_tri_table(void);   /* 256*16, -1 terminated */
const int *go_mc_pair_table(void);     /* 12*2 */

#ifdef __cplusplus
}
#endif
#endif
