/*
 * gsdf_program.h -- the packed SDF node program: the contract between the host-side flattener and the sm_100a
 * interpreter kernels.
 *
 * The reference walks its glbuild.Shader3D tree with ForEachChild / ForEach2DChild (glbuild/glbuild.go:63-89) and,
 * on its GPU path, turns every node into GLSL text (glbuild/glbuild.go:175-214).  Here the same walk emits a linear
 * postfix instruction stream that a stack machine executes once per point.  The stream has NO data-dependent control
 * flow: every thread of every warp executes the same opcode at the same time (warp-uniform dispatch).
 *
 * Machine state per point:   p = (x,y,z)  current position (2D nodes use x,y)
 *                            D            distance stack (top cached in a register)
 *                            P            position stack (only used where a later sibling needs the old p)
 *
 * Layout: the stream is an array of 16-byte chunks (4 x u32).  Chunk 0 of an instruction is the header
 *     w0 = opcode | (nchunks << 8)      nchunks counts the header itself
 *     w1, w2, w3 = op-specific raw 32-bit fields (int or float bits)
 * followed by nchunks-1 parameter chunks of 4 floats.  The side buffer `aux` holds variable-length float data
 * (polygon edge records, line segments); offsets into it are in floats.
 *
 * All derived constants (half sizes, reciprocals, tan(taper) ...) are computed by the flattener in float32 with the
 * same operation order cpu_evaluators.go uses per Evaluate call, so kernels stay bit-comparable with the CPU path.
 */
#ifndef GSDF_PROGRAM_H
#define GSDF_PROGRAM_H

#ifndef __CUDACC_RTC__ /* (run-time compiled kernels bring their own fixed-width types: no system headers under NVRTC) */
#include <stdint.h>
#endif

#define GSDF_PROGRAM_MAGIC 0x46445347u /* "GSDF" */
#define GSDF_PROGRAM_VERSION 1u

/* Opcode semantics. "push(v)": D.push(v). "top": D top. "below": pop the entry under the top.
 * Reference citations are cpu_evaluators.go unless a file is named. */
enum gsdf_opcode {
    GSDF_OP_END = 0,
    /* ---- 3D primitives: push(f(p)) ---- */
    GSDF_OP_SPHERE,      /* :20   w2=r */
    GSDF_OP_BOX,         /* :28   c1=(hx,hy,hz,round) */
    GSDF_OP_BOXFRAME,    /* :38   c1=(bx,by,bz,e) */
    GSDF_OP_TORUS,       /* :59   w2=rGreater w3=rLesser */
    GSDF_OP_CYLINDER,    /* :70   c1=(r,h',round,_) ; w1=1 when round!=0 */
    GSDF_OP_HEX,         /* :90   c1=(side,h,clm,_) */
    /* ---- 2D primitives: push(f(p.xy)) ---- */
    GSDF_OP_CIRCLE2D,    /* :661  w2=r */
    GSDF_OP_RECT2D,      /* :685  w2=bx w3=by */
    GSDF_OP_LINE2D,      /* :551  c1=(ax,ay,bax,bay) c2=(dotba,w,_,_) */
    GSDF_OP_LINES2D,     /* :1145 w1=aux_off w2=nseg w3=w ; aux: (ax,ay,bx,by) per segment */
    GSDF_OP_ARC2D,       /* :564  c1=(r,t,s,c) c2=(scrx,scry,_,_) */
    GSDF_OP_EQTRI2D,     /* :669  w2=r w3=r/k */
    GSDF_OP_HEX2D,       /* :718  w2=r w3=kz*r */
    GSDF_OP_OCT2D,       /* :731  w2=r w3=kz*r */
    GSDF_OP_DIAMOND2D,   /* :694  c1=(bx,by,dot(b,b),_) */
    GSDF_OP_ROUNDX2D,    /* :705  w2=w w3=r */
    GSDF_OP_POLY2D,      /* :793  w1=aux_off w2=nverts ; aux: (v1x,v1y,ex,ey,norm2e,v2y,_,_) per edge (8 floats) */
    GSDF_OP_ELLIPSE2D,   /* :750  w2=a w3=b */
    GSDF_OP_BEZIERQ2D,   /* :581  w2=thick/2 c1=(Ax,Ay,ax,ay) c2=(bx,by,cx,cy) c3=(kk,kx,kx2,a2) */
    /* ---- distance combiners: b=top, a=below, top=f(a,b) ---- */
    GSDF_OP_MIN,         /* :14,124,821  union fold */
    GSDF_OP_MAX,         /* :146,847     intersect */
    GSDF_OP_DIFF,        /* :168,869     max(a,-b) */
    GSDF_OP_XOR,         /* :190,891 */
    GSDF_OP_SMOOTH_UNION,     /* :213 w2=k */
    GSDF_OP_SMOOTH_DIFF,      /* :238 w2=k */
    GSDF_OP_SMOOTH_INTERSECT, /* :263 w2=k */
    /* ---- unary distance ops on top ---- */
    GSDF_OP_OFFSET,      /* :454,964  top += w2 */
    GSDF_OP_ANNULUS,     /* :1026     top = |top| - w2 */
    GSDF_OP_MULDIST,     /* :308,1222 top *= w2 (scale exit) */
    GSDF_OP_SHELL_EXIT,  /* :448      top = w2*(|top| - w2) */
    GSDF_OP_ADD_BELOW,   /* :422,1251 a=below; top = top + a (elongate exit) */
    GSDF_OP_EXTRUDE_EXIT,/* :524-529  wy=below; top = min(0,max(top,wy)) + hypot(max(top,0),max(wy,0)) */
    GSDF_OP_MAX_BELOW,   /* threads.go:176-180 a=below; top = max(top, a) (screw exit) */
    /* ---- position stack ---- */
    GSDF_OP_PUSH_POS,    /* P.push(p) */
    GSDF_OP_POP_POS,     /* p = P.pop() */
    GSDF_OP_PEEK_POS,    /* p = P.top() */
    /* ---- position transforms: p = f(p) ---- */
    GSDF_OP_TRANSLATE,   /* :470,980  c1=(tx,ty,tz,_) : p -= t (2D: tz=0) */
    GSDF_OP_SCALE_POS,   /* :300,437,1216  w2=inv : p *= inv */
    GSDF_OP_SYMMETRY,    /* :314,998  w1=mask bits x=1,y=2,z=4 */
    GSDF_OP_TRANSFORM,   /* :488      c1..c3 = rows of tInv (x00 x01 x02 x03 / ...) */
    GSDF_OP_ROTATE2D,    /* :1186     c1=(x00,x01,x10,x11) */
    GSDF_OP_TWIST,       /* :1257     w2=k */
    GSDF_OP_ELONGATE,    /* :399      c1=(hx,hy,hz,_) : q=|p|-h ; push(min(max(q),0)) ; p=max(q,0) */
    GSDF_OP_ELONGATE2D,  /* :1228     w2=hx w3=hy */
    GSDF_OP_ARRAY_VAR,   /* :345      w1=variant(i|j<<1|k<<2) c1=(sx,sy,sz,_) c2=(nx-1,ny-1,nz-1,_) */
    GSDF_OP_ARRAY2D_VAR, /* :914      w1=variant(i|j<<1) c1=(sx,sy,nx-1,ny-1) */
    GSDF_OP_CIRC_ENTER,  /* :1042,1094 c1=(angle,ncirc,ninsm1,_) : P.push(p0) ; p=p1 */
    GSDF_OP_EXTRUDE_ENTER, /* :506    w1=guard w2=h/2 w3=guard k : push(|z|-h/2) */
    GSDF_OP_REVOLVE,     /* :533      w2=off : p=(hypot(x,z)-off, y) */
    GSDF_OP_SCREW_ENTER, /* threads.go:141-170 w1=guard w3=guard k c1=(pitch,lead,L/2,tanTaper) : push(|z|-L/2) ; p=(saw,y) */
    /* ---- 2-D bounding-box culling of union / difference operands (see "box guards" below) ---- */
    GSDF_OP_CULL_UB2D,   /* w1=aux_off w2=npts w3=abs margin ; aux: anchor points (x,y) ON the operands' outlines.
                            push(U), U = (1+1e-4)*min_k |p - v_k| + margin >= min over the union's operands */
    GSDF_OP_BBOX_GUARD2D,/* w1=guard kind | target<<8, w3=abs margin, c1=(minx,miny,maxx,maxy) of the operand that follows:
                            w = (1-1e-4)*dist(p, box) - margin <= operand(p) for p OUTSIDE the box; the tile skips the operand iff every
                            point of it lies outside the box and satisfies guard_dead(kind, w, top) (a point inside the box
                            never votes for the skip: the box says nothing about the operand's value there)
                            => jump to `target` (the operand's combiner, which then keeps `top`) */
    GSDF_OP_MIN_CONST,   /* top = min(w2, top): the accumulator seed of the reference's array folds (largenum = 1e20,
                            cpu_evaluators.go:364,932; math.MaxFloat32 in translateMulti2D, :1172) applied behind the fold */
    GSDF_OP__COUNT
};

/* Box guards -- the slab-guard idea for 2-D unions of bounded shapes (glyphs of a text line, forge/textsdf/font.go:89-141).
 *
 * For operands built only from exact 2-D primitives (poly2D, circle, rect), translate2D, union2D and the minuend of
 * diff2D, the value at p is >= dist(p, Bounds()) outside the box and <= |p - v| for any point v on the operand's
 * outline. A union first pushes U (CULL_UB2D, from a few anchor points per operand) as an EXTRA operand of its min fold:
 * U >= the operand that owns the nearest anchor >= the true minimum, so min(U, d_1 .. d_n) == min(d_1 .. d_n) bit for
 * bit. Each operand is then preceded by BBOX_GUARD2D(MIN): when its box is farther than the running minimum for every
 * point of the CTA's tile, the operand (hundreds of polygon edges) is skipped and its MIN keeps the running value; the
 * operand holding the minimum can never be skipped because its lower bound is <= its value <= the running minimum.
 * The subtrahend of a diff2D is guarded the same way with GSDF_GUARD_DIFF (-w < a => max(a, -s) == a): the hole of a
 * glyph is only evaluated by tiles that reach into the glyph. Relative (1e-4) and absolute (1e-5 * coordinate scale)
 * margins absorb the float rounding of the bounds themselves. Image tiles are 128 x 16 pixels, so the predicate is
 * uniform over almost every tile. */

/* Slab guards -- the one place the stream has control flow, and it is CTA-uniform and value-preserving.
 *
 * An extrude or screw node returns a value s >= w = |z| - h/2 in its own frame (cpu_evaluators.go:524-529: the
 * extrusion formula is >= its w argument; threads.go:176-180: max(profile, w)).  w is known at the ENTER op, before
 * any of the 2-D work below it (a 100-edge polygon for a thread profile).  When that node is the later operand of a
 * combiner whose result cannot depend on any s >= w, the whole subtree is dead for that point:
 *     GSDF_GUARD_DIFF          a=top:  -w < a            =>  max(a, -s) == a
 *     GSDF_GUARD_MIN           a=top:   w > a            =>  min(a, s)  == a
 *     GSDF_GUARD_SMOOTH_UNION  a=top:   w - a >= k, a!=0 =>  h clamps to 1 and the blend returns a bit for bit
 * ENTER's w1 = kind | target<<8.  If the predicate holds for EVERY point of the CTA's tile (one __syncthreads_and) the
 * interpreter jumps to chunk `target` (just behind the node's exit op; restoring POP_POS ops still run) and the next
 * combiner is a no-op that leaves a on top.  Otherwise the tile runs the full stream.  Either way each point's result
 * is bit-identical to the unguarded program's.  Work tiles are runs of neighbouring lattice cells, so the predicate
 * is uniform over most tiles of a part whose threaded/extruded feature occupies a fraction of its volume. */
enum gsdf_guard_kind { GSDF_GUARD_NONE = 0, GSDF_GUARD_DIFF = 1, GSDF_GUARD_MIN = 2, GSDF_GUARD_SMOOTH_UNION = 3 };

/* Radius reuse. Part of the default build since round 2 (measured on B200: -5 % on the lattice evaluation of the
 * npt-flange, every GPU parity test bit-identical). The flags are optional hints: a flattener that never sets them emits
 * valid programs (each op then computes its own radius). GSDF_RXY=0 in the flattener's environment switches the post-pass
 * off (A/B); a library built with -DGSDF_NO_RXY has no radius slot and rejects flagged programs.
 *
 * CYLINDER, TORUS, CIRCLE2D and SCREW_ENTER all start from r = math32.Hypot(p.x, p.y): an IEEE division and square root,
 * about 30 instructions. On a part like the npt-flange the same x, y reach four such ops per evaluation (three coaxial
 * cylinders whose translations only move z, and the screw). A post-pass of the flattener tracks, through the straight-
 * line stream, which ops can change x or y (PUSH/POP/PEEK_POS restore an earlier state; a TRANSLATE by (+0, +0, tz) changes
 * nothing) and marks a consumer GSDF_RXY_READ when the one-slot radius cache provably holds Hypot of bit-identical x, y,
 * and GSDF_RXY_WRITE when it computes the radius for later readers. Writers are never inside a region a guard can skip,
 * so the cache content is static. Reusing a value computed from identical bits is bit-identical by construction.
 * Flag word: w1 for CYLINDER (bit 0 stays the rounding flag), TORUS and CIRCLE2D; w2 for SCREW_ENTER (w1 holds its guard). */
#ifndef GSDF_NO_RXY
#ifndef GSDF_RXY
#define GSDF_RXY 1
#endif
#endif
#define GSDF_RXY_READ 0x100u
#define GSDF_RXY_WRITE 0x200u

/* Header that precedes the instruction chunks in the blob handed to gsdf_program_create (one 16-byte chunk x2). */
typedef struct {
    uint32_t magic;      /* GSDF_PROGRAM_MAGIC */
    uint32_t version;    /* GSDF_PROGRAM_VERSION */
    uint32_t nchunks;    /* number of 16-byte instruction chunks that follow (END included) */
    uint32_t dim;        /* 3 or 2: dimension of the root */
    uint32_t dstack;     /* distance-stack slots needed below the cached top */
    uint32_t pstack;     /* position-stack slots needed */
    uint32_t ninstr;     /* instruction count (informational) */
    uint32_t reserved;
} gsdf_program_header;   /* 32 bytes */

#define GSDF_POLY_EDGE_FLOATS 8

#endif
