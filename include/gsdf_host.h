/*
 * gsdf_host.h -- flat C wrapper over the C++ host layer (gsdf_b200/csrc/host): the mirror of gsdf.Builder,
 * forge/threads and the flattener, so Python (ctypes) tests and bench.py can build the reference's scenes without a
 * Go toolchain.  In a Go deployment this layer is NOT used: the Go flattener emits the node program
 * directly from the live glbuild.Shader3D tree and calls include/gsdf_b200.h.
 */
#ifndef GSDF_HOST_H
#define GSDF_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "gsdf_b200.h" /* status codes only: libgsdfhost.so calls nothing of libgsdfb200.so */
#include "gsdf_tree.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsdfh_builder gsdfh_builder;

gsdfh_builder *gsdfh_builder_new(void);
void gsdfh_builder_free(gsdfh_builder *b);
/* gsdf.Builder.Err() (gsdf.go:88): accumulated shape errors joined by '\n', "" when none. */
const char *gsdfh_builder_err(gsdfh_builder *b);
void gsdfh_builder_clear_errors(gsdfh_builder *b);

/* Extra call codes for gsdfh_node beyond gsdf_node_kind (constructors that are not 1:1 with a node type). */
enum {
    GSDFH_CALL_ROTATE = 200,        /* Builder.Rotate(s, radians, axis): f = {radians, ax, ay, az} (operations.go:394) */
    GSDFH_CALL_TRANSFORM16 = 201,   /* Builder.Transform(s, m): f = 16 floats row-major (operations.go:340) */
    GSDFH_CALL_TRIPRISM = 202,      /* NewTriangularPrism(triHeight, extrudeLength) (primitives.go:198) */
    GSDFH_CALL_BOUNDSBOXFRAME = 203 /* NewBoundsBoxFrame(bb): f = {min.xyz, max.xyz} (primitives.go:12) */
};
/* Generic constructor: `kind` is a gsdf_node_kind (dispatches to the Builder method of the same name, with its
 * validation) or a GSDFH_CALL_* code. f/ip are the constructor's float/int arguments in declaration order;
 * children are node ids; aux is variable-length float data (polygon vertices, segments, displacements).
 * Returns the new node id, or <0 (message via gsdfh_builder_err). */
int32_t gsdfh_node(gsdfh_builder *b, int32_t kind, const float *f, int nf, const int32_t *ip, int ni, const int32_t *children,
                   int nchild, const float *aux, int naux);

int gsdfh_is2d(gsdfh_builder *b, int32_t id);
/* Shader3D.Bounds() / Shader2D.Bounds(): out = {min..., max...} */
int gsdfh_bounds3(gsdfh_builder *b, int32_t id, float out[6]);
int gsdfh_bounds2(gsdfh_builder *b, int32_t id, float out[4]);

/* forge/threads front-ends. thread_kind: 0 ISO{D,P,Ext} (iso.go:20), 1 NPT from nominal size in p0 (npt.go:63),
 * 2 UTS{D,TPI,Ext} (uts.go:8), 3 Acme{D,P} (acme.go:10), 4 ANSIButtress{D,P} (ansibuttress.go:10), 5 PlasticButtress{D,P}
 * (plasticbuttress.go:9). */
int32_t gsdfh_thread_profile(gsdfh_builder *b, int thread_kind, float p0, float p1, int ext);                 /* Threader.Thread */
int32_t gsdfh_screw(gsdfh_builder *b, float length, int thread_kind, float p0, float p1, int ext);            /* threads.Screw */
int32_t gsdfh_nut(gsdfh_builder *b, int thread_kind, float p0, float p1, int ext, int style, float tol);      /* threads.Nut */
int32_t gsdfh_bolt(gsdfh_builder *b, int thread_kind, float p0, float p1, int ext, int style, float tol, float total_len,
                   float shank_len);                                                                          /* threads.Bolt */
int32_t gsdfh_hexhead(gsdfh_builder *b, float radius, float height, int round_neg, int round_pos);            /* threads.HexHead */
/* The example scenes BASELINE.json names: "npt-flange", "bolt", "knurled-cylinder" (param = diameter, 0 -> 20); plus
 * "fibonacci-showerhead" (the reference README's second timed example, a known-answer test for the oracle). */
int32_t gsdfh_scene(gsdfh_builder *b, const char *name, float param);

/* forge/textsdf (font.go): TrueType font -> polygon SDF tree. The font bytes are supplied by the caller (the reference
 * embeds iso-3098.ttf, embed.go:10-16; this library ships no font). */
typedef struct gsdfh_font gsdfh_font;
gsdfh_font *gsdfh_font_new(void);
void gsdfh_font_free(gsdfh_font *f);
const char *gsdfh_font_err(gsdfh_font *f);
/* Font.Configure(FontConfig{RelativeGlyphTolerance}) (font.go:40-51); 0 = default. Returns 0 or <0. */
int gsdfh_font_configure(gsdfh_font *f, float relative_glyph_tolerance);
/* Font.LoadTTFBytes (font.go:54-62). */
int gsdfh_font_load_ttf(gsdfh_font *f, const void *ttf, size_t nbytes);
/* Font.TextLine(s) (font.go:89-141): utf-8 text -> root node id in builder b, or <0 (gsdfh_font_err). */
int32_t gsdfh_font_textline(gsdfh_font *f, gsdfh_builder *b, const char *utf8);
/* Font.Glyph(c) (font.go:159-165). */
int32_t gsdfh_font_glyph(gsdfh_font *f, gsdfh_builder *b, uint32_t rune);
/* Font.Kern / Font.AdvanceWidth (font.go:144-156) and the unexported scaleout (font.go:208-212). */
float gsdfh_font_kern(gsdfh_font *f, uint32_t c0, uint32_t c1);
float gsdfh_font_advance_width(gsdfh_font *f, uint32_t c);
float gsdfh_font_scaleout(gsdfh_font *f);
/* sfnt pieces the reference calls, for parser tests: glyph index of a rune; contour segments of a glyph at
 * ppem = unitsPerEm as rows {op, x0,y0, x1,y1, x2,y2} (op 0 MoveTo, 1 LineTo, 2 QuadTo, 3 CubeTo; Y down). Returns the
 * number of segments (call with out==NULL to size), <0 on error. info = {unitsPerEm, numGlyphs, bounds min.x, min.y,
 * max.x, max.y}. */
int32_t gsdfh_font_glyph_index(gsdfh_font *f, uint32_t rune);
int32_t gsdfh_font_glyph_segments(gsdfh_font *f, int32_t glyph_index, int32_t *out7, int32_t max_segments);
int gsdfh_font_info(gsdfh_font *f, int32_t info[6]);

/* Tree table export (for the CPU oracle in tests). Pointers stay valid until the builder is next mutated. */
int gsdfh_tree(gsdfh_builder *b, const gsdf_tree_node **nodes, int32_t *nnodes, const int32_t **children, int32_t *nchildren,
               const float **aux, int32_t *naux);

/* Flattener: tree -> node program (include/gsdf_program.h). */
typedef struct gsdfh_flat gsdfh_flat;
gsdfh_flat *gsdfh_flatten(gsdfh_builder *b, int32_t root);
const void *gsdfh_flat_blob(const gsdfh_flat *f, size_t *nbytes);
const float *gsdfh_flat_aux(const gsdfh_flat *f, size_t *nfloats);
/* info = {dim, ninstr, nchunks, dstack, pstack} */
void gsdfh_flat_info(const gsdfh_flat *f, int32_t info[5]);
void gsdfh_flat_free(gsdfh_flat *f);

/* NewCUDASDF3 / NewCUDASDF2 (mirrors gleval.NewComputeGPUSDF3, gleval/gpu.go:35) = gsdfh_flatten here + gsdf_program_create of
 * libgsdfb200.so on the blob: this library has no CUDA dependency (bench.py's CPU reference arm loads it alone). */

#ifdef __cplusplus
}
#endif
#endif
