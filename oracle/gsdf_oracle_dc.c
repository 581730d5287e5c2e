/*
 * gsdf_oracle_dc.c -- CPU ORACLE, dual contouring part.  TEST INFRASTRUCTURE ONLY (see gsdf_oracle.h).
 *
 * Restates glrender/dual_contour.go (DualContourRenderer.Reset / RenderAll, DualCube), glrender/
 * dual_contour_vertexplacement.go (DualContourLeastSquares, leastSquaresMGS64), the DualContourNaive placer of
 * glrender/dual_contour_test.go:355-389 and gleval.NormalsCentralDiff (gleval/gleval.go:53-108).
 *
 * PARITY: the reference's tests hold no golden vertex arrays for this path, only properties (vertices within 1.5 res
 * of the surface, least squares no worse than the naive mean, no vertex inside the shape / outside the bounds:
 * dual_contour_test.go:140-497); tests/test_dual_contour.py restates those. UNPINNED: ms3.Octree.DecomposeBFS cube
 * order (assumed: level by level, children in i3.Cube.Octree() = Bourke corner order, so the level-1 cubes come out
 * ordered by their per-level child indices, most significant level first), ms3.Octree.CubeOrigin / CubeSize,
 * ms3.Box.Add (translation), ms3.Dot / ClampElem -- all from the un-vendored github.com/soypat/geometry module.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "gsdf_oracle.h"

typedef struct { float x, y, z; } v3;

/* per-level child index of i3.Cube.Octree(): (0,0,0)(1,0,0)(1,1,0)(0,1,0)(0,0,1)(1,0,1)(1,1,1)(0,1,1) */
static uint64_t dc_key(int i, int j, int k, int bits) {
    uint64_t key = 0;
    for (int b = bits - 1; b >= 0; b--) {
        int xb = (i >> b) & 1, yb = (j >> b) & 1, zb = (k >> b) & 1;
        int c = zb * 4 + (yb ? 3 - xb : xb);
        key = (key << 3) | (uint64_t)c;
    }
    return key;
}
static void dc_unkey(uint64_t key, int bits, int *i, int *j, int *k) {
    int x = 0, y = 0, z = 0;
    for (int b = bits - 1; b >= 0; b--) {
        int c = (int)((key >> (3 * b)) & 7);
        int zb = c >> 2, r = c & 3, yb = r >> 1, xb = yb ? 3 - r : r;
        x |= xb << b; y |= yb << b; z |= zb << b;
    }
    *i = x; *j = y; *k = z;
}

/* makeICube (glrender/octreerenderer.go:222-235) on Bounds().Add(-res/2) (dual_contour.go:31-34). */
int go_dc_levels(const float bbmin[3], const float bbmax[3], float res, float origin[3]) {
    if (!(res > 0) || isinf(res)) return -1;
    float sub = res / 2;
    float mn[3], mx[3];
    for (int a = 0; a < 3; a++) { mn[a] = bbmin[a] + -sub; mx[a] = bbmax[a] + -sub; }
    float longAxis = fmaxf(mx[0] - mn[0], fmaxf(mx[1] - mn[1], mx[2] - mn[2]));
    int levels = (int)ceilf(log2f(longAxis / res)) + 1;
    if (levels <= 1) return -1;
    if (origin) { origin[0] = mn[0]; origin[1] = mn[1]; origin[2] = mn[2]; }
    return levels;
}

typedef struct {
    int i, j, k;
    float o, xd, yd, zd; /* DualCube.OrigDist, XDist, YDist, ZDist */
    v3 fin;              /* FinalVertex */
    int nnb;             /* len(Neighbors) */
    int nb[12][2];       /* {cube index, axis}, in append order */
} dcube;

static int actx(const dcube *c) { return signbit(c->o) != signbit(c->xd); } /* dual_contour.go:266-274 */
static int acty(const dcube *c) { return signbit(c->o) != signbit(c->yd); }
static int actz(const dcube *c) { return signbit(c->o) != signbit(c->zd); }
static float isx(const dcube *c) { return -c->o / (c->xd - c->o); }         /* :275-277 */
static float isy(const dcube *c) { return -c->o / (c->yd - c->o); }
static float isz(const dcube *c) { return -c->o / (c->zd - c->o); }
static v3 corigin(const float org[3], float res, const dcube *c) {           /* ms3.Octree.CubeOrigin, level 1 */
    v3 p = {org[0] + res * (float)c->i, org[1] + res * (float)c->j, org[2] + res * (float)c->k};
    return p;
}
/* EdgeNeighborsX/Y/Z (dual_contour.go:282-298) in cell units */
static const int ENB[3][4][3] = {
    {{0, -1, -1}, {0, 0, -1}, {0, 0, 0}, {0, -1, 0}},
    {{-1, 0, -1}, {-1, 0, 0}, {0, 0, 0}, {0, 0, -1}},
    {{-1, -1, 0}, {0, -1, 0}, {0, 0, 0}, {-1, 0, 0}},
};

/* leastSquaresMGS64 (dual_contour_vertexplacement.go:148-223) */
static void lsq_mgs64(int K, float A[][3], const float *b, float x3[3]) {
    x3[0] = x3[1] = x3[2] = 0;
    if (K < 3) return;
    double Q[32][3], b64[32], R[3][3] = {{0}};
    for (int k = 0; k < K; k++) { for (int c = 0; c < 3; c++) Q[k][c] = (double)A[k][c]; b64[k] = (double)b[k]; }
    for (int j = 0; j < 3; j++) {
        for (int i = 0; i < j; i++) {
            double dot = 0;
            for (int k = 0; k < K; k++) dot += Q[k][i] * Q[k][j];
            R[i][j] = dot;
            for (int k = 0; k < K; k++) Q[k][j] -= dot * Q[k][i];
        }
        double normSq = 0;
        for (int k = 0; k < K; k++) normSq += Q[k][j] * Q[k][j];
        double norm = sqrt(normSq);
        R[j][j] = norm;
        if (norm > 1e-14) { double inv = 1.0 / norm; for (int k = 0; k < K; k++) Q[k][j] *= inv; }
    }
    double Qtb[3] = {0, 0, 0};
    for (int j = 0; j < 3; j++) for (int k = 0; k < K; k++) Qtb[j] += Q[k][j] * b64[k];
    double x[3];
    for (int i = 2; i >= 0; i--) {
        x[i] = Qtb[i];
        for (int k = i + 1; k < 3; k++) x[i] -= R[i][k] * x[k];
        if (R[i][i] > 1e-14) x[i] /= R[i][i]; else x[i] = 0;
    }
    for (int c = 0; c < 3; c++) x3[c] = (float)x[c];
}
/* exported for the known-answer tests (dual_contour_test.go:20-137 state two QEF systems with known solutions) */
int go_lsq_mgs64(int K, const float *A_rowmajor, const float *b, float x3[3]) {
    if (K < 0 || K > 32) return -1;
    float A[32][3];
    for (int k = 0; k < K; k++) for (int c = 0; c < 3; c++) A[k][c] = A_rowmajor[3 * k + c];
    lsq_mgs64(K, A, b, x3);
    return 0;
}
static float clampf_(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* placer: 0 = DualContourNaive, 1 = DualContourLeastSquares{}, 2 = DualContourLeastSquares{Chiseled: true}.
 * tri9: up to max_tris triangles; returns the triangle count (counting continues past max_tris) or <0.
 * stats (optional, 4 x int64): {levels, cubes kept by the prune, cubes with >= 1 neighbour entry, SDF evaluations}. */
int64_t go_dual_contour_ex(const go_tree *t, const float bbmin[3], const float bbmax[3], float res, int placer, float *tri9, int64_t max_tris,
                           int64_t *stats, uint32_t *quad_keys);
int64_t go_dual_contour(const go_tree *t, const float bbmin[3], const float bbmax[3], float res, int placer, float *tri9, int64_t max_tris,
                        int64_t *stats) {
    return go_dual_contour_ex(t, bbmin, bbmax, res, placer, tri9, max_tris, stats, NULL);
}
/* quad_keys (optional, one per quad = per two triangles, same max_tris bound): BFS key of the cube that emitted the quad. */
int64_t go_dual_contour_ex(const go_tree *t, const float bbmin[3], const float bbmax[3], float res, int placer, float *tri9, int64_t max_tris,
                           int64_t *stats, uint32_t *quad_keys) {
    float org[3];
    int levels = go_dc_levels(bbmin, bbmax, res, org);
    if (levels <= 1) return -1;
    if (levels > 9) return -3; /* the oracle keeps dense N^3 tables: N <= 256 */
    const int bits = levels - 1, N = 1 << bits;
    const uint64_t ncell = (uint64_t)N * N * N;
    int64_t evals = 0;
    /* Reset: every level-1 cube, pruned at its ORIGIN with szMultMaxDist 2 (dual_contour.go:57, octreerenderer.go:240-284) */
    v3 *pos = (v3 *)malloc(sizeof(v3) * ncell);
    float *dist = (float *)malloc(sizeof(float) * ncell);
    int32_t *map = (int32_t *)malloc(sizeof(int32_t) * ncell); /* cubeMap: cell (x fastest) -> cube index */
    if (!pos || !dist || !map) return -2;
    for (uint64_t key = 0; key < ncell; key++) {
        int i, j, k;
        dc_unkey(key, bits, &i, &j, &k);
        pos[key].x = org[0] + res * (float)i; pos[key].y = org[1] + res * (float)j; pos[key].z = org[2] + res * (float)k;
    }
    if (go_eval3(t, (const float *)pos, dist, ncell)) return -2;
    evals += (int64_t)ncell;
    size_t nc = 0;
    for (uint64_t key = 0; key < ncell; key++) if (!(fabsf(dist[key]) >= res * 2)) nc++;
    dcube *cubes = (dcube *)calloc(nc ? nc : 1, sizeof(dcube));
    memset(map, 0xff, sizeof(int32_t) * ncell);
    size_t e = 0;
    for (uint64_t key = 0; key < ncell; key++) {
        if (fabsf(dist[key]) >= res * 2) continue;
        dcube *c = &cubes[e];
        dc_unkey(key, bits, &c->i, &c->j, &c->k);
        map[((size_t)c->k * N + c->j) * N + c->i] = (int32_t)e;
        e++;
    }
    free(pos); free(dist);
    /* RenderAll: origin and the three edge ends of every cube (dual_contour.go:85-107) */
    v3 *p4 = (v3 *)malloc(sizeof(v3) * 4 * (nc ? nc : 1));
    float *d4 = (float *)malloc(sizeof(float) * 4 * (nc ? nc : 1));
    for (size_t c = 0; c < nc; c++) {
        v3 o = corigin(org, res, &cubes[c]);
        p4[4 * c] = o;
        p4[4 * c + 1] = (v3){o.x + res, o.y + 0.f, o.z + 0.f};
        p4[4 * c + 2] = (v3){o.x + 0.f, o.y + res, o.z + 0.f};
        p4[4 * c + 3] = (v3){o.x + 0.f, o.y + 0.f, o.z + res};
    }
    if (nc && go_eval3(t, (const float *)p4, d4, 4 * nc)) return -2;
    evals += 4 * (int64_t)nc;
    for (size_t c = 0; c < nc; c++) {
        cubes[c].o = d4[4 * c]; cubes[c].xd = d4[4 * c + 1]; cubes[c].yd = d4[4 * c + 2]; cubes[c].zd = d4[4 * c + 3];
        cubes[c].fin = corigin(org, res, &cubes[c]); /* default FinalVertex (:114) */
    }
    free(p4); free(d4);
#define CELL(i_, j_, k_) (((i_) < 0 || (j_) < 0 || (k_) < 0 || (i_) >= N || (j_) >= N || (k_) >= N) ? -1 : map[((size_t)(k_) * N + (j_)) * N + (i_)])
    /* second loop: accumulate edge neighbours (dual_contour.go:118-146) */
    for (size_t c = 0; c < nc; c++) {
        const dcube *cu = &cubes[c];
        int act[3] = {actx(cu), acty(cu), actz(cu)};
        for (int a = 0; a < 3; a++) {
            if (!act[a]) continue;
            for (int q = 0; q < 4; q++) {
                int idx = CELL(cu->i + ENB[a][q][0], cu->j + ENB[a][q][1], cu->k + ENB[a][q][2]);
                if (idx >= 0) { dcube *n = &cubes[idx]; n->nb[n->nnb][0] = (int)c; n->nb[n->nnb][1] = a; n->nnb++; }
            }
        }
    }
    int64_t withnb = 0;
    for (size_t c = 0; c < nc; c++) withnb += cubes[c].nnb > 0;
    /* PlaceVertices */
    if (placer == 0) { /* dual_contour_test.go:358-389 */
        for (size_t c = 0; c < nc; c++) {
            dcube *cu = &cubes[c];
            if (cu->nnb == 0) continue;
            v3 sum = {0, 0, 0};
            for (int n = 0; n < cu->nnb; n++) {
                const dcube *nb = &cubes[cu->nb[n][0]];
                v3 o = corigin(org, res, nb), ct = o;
                if (cu->nb[n][1] == 0) ct = (v3){o.x + res * isx(nb), o.y + 0.f, o.z + 0.f};
                else if (cu->nb[n][1] == 1) ct = (v3){o.x + 0.f, o.y + res * isy(nb), o.z + 0.f};
                else ct = (v3){o.x + 0.f, o.y + 0.f, o.z + res * isz(nb)};
                sum.x += ct.x; sum.y += ct.y; sum.z += ct.z;
            }
            float inv = 1.0f / (float)cu->nnb;
            cu->fin = (v3){sum.x * inv, sum.y * inv, sum.z * inv};
        }
    } else {
        /* normals at the three edge intersections of every cube (dual_contour_vertexplacement.go:28-50) */
        const double normStep = placer == 2 ? 1e-4 : 2e-8;
        float step = (float)normStep;
        step *= 0.5f; /* gleval.go:54 */
        size_t np = 3 * nc;
        v3 *ip = (v3 *)malloc(sizeof(v3) * (np ? np : 1)), *aux = (v3 *)malloc(sizeof(v3) * (np ? np : 1));
        v3 *nrm = (v3 *)calloc(np ? np : 1, sizeof(v3));
        float *d1 = (float *)malloc(sizeof(float) * (np ? np : 1)), *d2 = (float *)malloc(sizeof(float) * (np ? np : 1));
        for (size_t c = 0; c < nc; c++) {
            const dcube *cu = &cubes[c];
            v3 o = corigin(org, res, cu);
            ip[3 * c] = (v3){o.x + res * isx(cu), o.y + 0.f, o.z + 0.f};
            ip[3 * c + 1] = (v3){o.x + 0.f, o.y + res * isy(cu), o.z + 0.f};
            ip[3 * c + 2] = (v3){o.x + 0.f, o.y + 0.f, o.z + res * isz(cu)};
        }
        for (int dim = 0; dim < 3 && np; dim++) { /* gleval.go:73-106 */
            float h[3] = {0, 0, 0};
            h[dim] = step;
            for (size_t q = 0; q < np; q++) aux[q] = (v3){ip[q].x + h[0], ip[q].y + h[1], ip[q].z + h[2]};
            if (go_eval3(t, (const float *)aux, d1, np)) return -2;
            for (size_t q = 0; q < np; q++) aux[q] = (v3){ip[q].x - h[0], ip[q].y - h[1], ip[q].z - h[2]};
            if (go_eval3(t, (const float *)aux, d2, np)) return -2;
            evals += 2 * (int64_t)np;
            for (size_t q = 0; q < np; q++) {
                float v = d1[q] - d2[q];
                if (dim == 0) nrm[q].x = v; else if (dim == 1) nrm[q].y = v; else nrm[q].z = v;
            }
        }
        const float invRes = 1.0f / res;
        const float sqrtLambda = placer == 2 ? (float)(sqrt(1e-5) * normStep) : (float)sqrt(1e-5);
        for (size_t c = 0; c < nc; c++) {
            dcube *cu = &cubes[c];
            if (cu->nnb == 0) continue;
            v3 co = corigin(org, res, cu);
            v3 bv[16], ln[16];
            int nb = 0;
            if (actx(cu)) { bv[nb] = (v3){co.x + res * isx(cu), co.y + 0.f, co.z + 0.f}; ln[nb++] = nrm[3 * c]; }
            if (acty(cu)) { bv[nb] = (v3){co.x + 0.f, co.y + res * isy(cu), co.z + 0.f}; ln[nb++] = nrm[3 * c + 1]; }
            if (actz(cu)) { bv[nb] = (v3){co.x + 0.f, co.y + 0.f, co.z + res * isz(cu)}; ln[nb++] = nrm[3 * c + 2]; }
            for (int n = 0; n < cu->nnb; n++) {
                const dcube *q = &cubes[cu->nb[n][0]];
                int axis = cu->nb[n][1];
                v3 o = corigin(org, res, q), ct;
                if (axis == 0) ct = (v3){o.x + res * isx(q), o.y + 0.f, o.z + 0.f};
                else if (axis == 1) ct = (v3){o.x + 0.f, o.y + res * isy(q), o.z + 0.f};
                else ct = (v3){o.x + 0.f, o.y + 0.f, o.z + res * isz(q)};
                bv[nb] = ct;
                ln[nb++] = nrm[3 * (size_t)cu->nb[n][0] + axis];
            }
            float A[32][3], b[32];
            int K = 0;
            v3 mean = {0, 0, 0};
            for (int r = 0; r < nb; r++) {
                v3 qi = {invRes * (bv[r].x - co.x), invRes * (bv[r].y - co.y), invRes * (bv[r].z - co.z)};
                A[K][0] = ln[r].x; A[K][1] = ln[r].y; A[K][2] = ln[r].z;
                b[K] = ln[r].x * qi.x + ln[r].y * qi.y + ln[r].z * qi.z;
                K++;
                mean.x += bv[r].x; mean.y += bv[r].y; mean.z += bv[r].z;
            }
            float invn = 1.f / (float)nb;
            mean = (v3){invn * mean.x, invn * mean.y, invn * mean.z};
            v3 bias = {invRes * (mean.x - co.x), invRes * (mean.y - co.y), invRes * (mean.z - co.z)};
            A[K][0] = sqrtLambda; A[K][1] = 0; A[K][2] = 0; b[K++] = sqrtLambda * bias.x;
            A[K][0] = 0; A[K][1] = sqrtLambda; A[K][2] = 0; b[K++] = sqrtLambda * bias.y;
            A[K][0] = 0; A[K][1] = 0; A[K][2] = sqrtLambda; b[K++] = sqrtLambda * bias.z;
            float x[3];
            lsq_mgs64(K, A, b, x);
            for (int a = 0; a < 3; a++) x[a] = clampf_(x[a], -0.1f, 1.1f);
            cu->fin = (v3){res * x[0] + co.x, res * x[1] + co.y, res * x[2] + co.z};
        }
        free(ip); free(aux); free(nrm); free(d1); free(d2);
    }
    /* quads -> triangles (dual_contour.go:152-218) */
    int64_t nt = 0;
    for (size_t c = 0; c < nc; c++) {
        const dcube *cu = &cubes[c];
        int act[3] = {actx(cu), acty(cu), actz(cu)};
        int flip[3] = {cu->xd - cu->o < 0, cu->yd - cu->o < 0, cu->zd - cu->o < 0};
        for (int a = 0; a < 3; a++) {
            if (!act[a]) continue;
            v3 quad[4];
            int all = 1;
            for (int q = 0; q < 4; q++) {
                int idx = CELL(cu->i + ENB[a][q][0], cu->j + ENB[a][q][1], cu->k + ENB[a][q][2]);
                if (idx < 0) { all = 0; break; }
                quad[q] = cubes[idx].fin;
            }
            if (!all) continue;
            if (flip[a]) { v3 t0 = quad[0], t1 = quad[1]; quad[0] = quad[3]; quad[1] = quad[2]; quad[2] = t1; quad[3] = t0; }
            const v3 tri[2][3] = {{quad[0], quad[1], quad[2]}, {quad[2], quad[3], quad[0]}};
            if (quad_keys && nt + 1 < max_tris) quad_keys[nt / 2] = (uint32_t)dc_key(cu->i, cu->j, cu->k, bits);
            for (int w = 0; w < 2; w++) {
                if (tri9 && nt < max_tris) memcpy(tri9 + 9 * nt, tri[w], 36);
                nt++;
            }
        }
    }
#undef CELL
    if (stats) { stats[0] = levels; stats[1] = (int64_t)nc; stats[2] = withnb; stats[3] = evals; }
    free(cubes); free(map);
    return nt;
}
