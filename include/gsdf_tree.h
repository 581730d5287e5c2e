/*
 * gsdf_tree.h -- the CSG tree table: one record per node holding the raw fields of the reference's Go node structs
 * (primitives.go, operations.go, primitives2d.go, operations2d.go, forge/threads/threads.go:62-69).
 *
 * It is what the host-side Builder stores, what the flattener walks to emit a node program
 * (include/gsdf_program.h), and what tests hand to the CPU oracle.  Children are listed in ForEachChild /
 * ForEach2DChild order (glbuild/glbuild.go:63-89).
 */
#ifndef GSDF_TREE_H
#define GSDF_TREE_H

#include <stdint.h>

enum gsdf_node_kind {
    /* 3D primitives (primitives.go)        fparam layout                                   */
    GSDF_N_SPHERE = 1,   /* :23   r                                                       */
    GSDF_N_BOX,          /* :75   dims.xyz, round                                         */
    GSDF_N_CYLINDER,     /* :119  r, h, round                                             */
    GSDF_N_HEX,          /* :164  side, h                                                 */
    GSDF_N_TORUS,        /* :207  rLesser, rGreater                                       */
    GSDF_N_BOXFRAME,     /* :266  dims.xyz, e                                             */
    /* 3D operations (operations.go) */
    GSDF_N_UNION = 16,   /* :27   n children                                              */
    GSDF_N_DIFF,         /* :124                                                          */
    GSDF_N_INTERSECT,    /* :167                                                          */
    GSDF_N_XOR,          /* :212                                                          */
    GSDF_N_SMOOTH_UNION, /* :570  k                                                       */
    GSDF_N_SMOOTH_DIFF,  /* :618  k                                                       */
    GSDF_N_SMOOTH_INTERSECT, /* :650 k                                                    */
    GSDF_N_SCALE,        /* :252  scale                                                   */
    GSDF_N_SYMMETRY,     /* :292  iparam[0] = x|y<<1|z<<2                                 */
    GSDF_N_TRANSFORM,    /* :350  tInv rows 0..2 (12 floats, row-major) ; aux = t rows 0..2 (for Bounds) */
    GSDF_N_TRANSLATE,    /* :407  p.xyz                                                   */
    GSDF_N_OFFSET,       /* :450  off                                                     */
    GSDF_N_ARRAY,        /* :498  d.xyz ; iparam nx,ny,nz                                 */
    GSDF_N_ELONGATE,     /* :683  h.xyz                                                   */
    GSDF_N_SHELL,        /* :727  thick                                                   */
    GSDF_N_CIRCARRAY,    /* :777  iparam nInst, circleDiv                                 */
    GSDF_N_TWIST,        /* :845  k                                                       */
    GSDF_N_BOUNDS3,      /* glbuild/glbuild.go:1080-1102 overloadBounds3: bb min.xyz, max.xyz ; Evaluate forwards */
    /* 2D -> 3D */
    GSDF_N_EXTRUDE = 40, /* operations2d.go:114  h                                        */
    GSDF_N_REVOLVE,      /* operations2d.go:163  off                                      */
    GSDF_N_SCREW,        /* forge/threads/threads.go:62  pitch, lead, lengthDiv2, taper   */
    /* 2D primitives (primitives2d.go) */
    GSDF_N_LINE2D = 64,  /* :33   width, a.xy, b.xy                                       */
    GSDF_N_LINES2D,      /* :92   width ; aux = (ax,ay,bx,by) per segment                 */
    GSDF_N_ARC2D,        /* :189  radius, angle, thick                                    */
    GSDF_N_CIRCLE2D,     /* :223  r                                                       */
    GSDF_N_EQTRI2D,      /* :261  hTri                                                    */
    GSDF_N_RECT2D,       /* :303  d.xy                                                    */
    GSDF_N_HEX2D,        /* :344  side                                                    */
    GSDF_N_OCT2D,        /* :381  c                                                       */
    GSDF_N_ELLIPSE2D,    /* :417  a, b                                                    */
    GSDF_N_POLY2D,       /* :454  aux = (x,y) per vertex                                  */
    GSDF_N_DIAMOND2D,    /* :556  d.xy                                                    */
    GSDF_N_ROUNDX2D,     /* :597  dim, thick                                              */
    GSDF_N_BEZIERQ2D,    /* :637  a.xy, b.xy, c.xy, thick                                 */
    /* 2D operations (operations2d.go) */
    GSDF_N_UNION2D = 96, /* :15                                                           */
    GSDF_N_DIFF2D,       /* :209                                                          */
    GSDF_N_INTERSECT2D,  /* :253                                                          */
    GSDF_N_XOR2D,        /* :297                                                          */
    GSDF_N_ARRAY2D,      /* :343  d.xy ; iparam nx, ny                                    */
    GSDF_N_OFFSET2D,     /* :416  f                                                       */
    GSDF_N_TRANSLATE2D,  /* :461  p.xy                                                    */
    GSDF_N_ROTATE2D,     /* :508  tInv (x00,x01,x10,x11), t (x00,x01,x10,x11)             */
    GSDF_N_SYMMETRY2D,   /* :562  iparam[0] = x|y<<1                                      */
    GSDF_N_ANNULUS2D,    /* :616  r                                                       */
    GSDF_N_CIRCARRAY2D,  /* :668  iparam nInst, circleDiv                                 */
    GSDF_N_SCALE2D,      /* :723  scale                                                   */
    GSDF_N_TRANSLATEMULTI2D, /* :767 aux = (x,y) per displacement                         */
    GSDF_N_ELONGATE2D,   /* :816  h.xy                                                    */
    GSDF_N_BOUNDS2       /* glbuild/glbuild.go:1105-1128 overloadBounds2: bb min.xy, max.xy ; Evaluate forwards */
};

typedef struct {
    int32_t kind;       /* gsdf_node_kind */
    int32_t nchild;
    int32_t child_off;  /* index of the first child id in the children array */
    int32_t aux_off;    /* float offset into the aux array */
    int32_t aux_cnt;    /* floats */
    int32_t iparam[3];
    float   fparam[16];
} gsdf_tree_node;       /* 96 bytes; same layout the oracle reads */

#endif
