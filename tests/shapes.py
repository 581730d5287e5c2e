"""Scene lists for the parity tests. They restate the reference's own test corpora:
   testPrimitives3D (gsdf_test.go:182-201), testBinOp3D (:203-231), testRandomUnary3D (:255-283),
   testPrimitives2D (:285-353), testBinary2D (:355-373), testRandomUnary2D (:233-253),
   examples/test/glsdf3test.go:100-114 (screw / NPT profile).
Go's math/rand stream (rand.NewSource(1), gsdf_test.go:49) cannot be regenerated without Go, so the randomised
unary operations list explicit parameter draws inside the generators' ranges (gsdf_test.go:572-730)."""
import math

import numpy as np

from gsdf_b200 import gsdf


def primitives3d(bld):
    maxdim = 1.0
    dx, dy, dz = maxdim, maxdim * 0.47, maxdim * 0.8
    thick = maxdim / 10
    return [
        ("sphere", bld.NewSphere(1)),
        ("box", bld.NewBox(dx, dy, dz, thick)),
        ("boxframe", bld.NewBoxFrame(dx, dy, dz, thick)),
        ("cylinder", bld.NewCylinder(dx, dy, 0)),
        ("cylinder_round", bld.NewCylinder(dx, dy, thick)),
        ("hexprism", bld.NewHexagonalPrism(dx, dy)),
        ("torus", bld.NewTorus(dx, dy)),
        ("triprism", bld.NewTriangularPrism(1, 0.5)),
    ]


def binops3d(bld):
    s1 = bld.NewSphere(1)
    s2 = bld.Translate(bld.NewBox(1, 0.6, .8, 0.1), 0.5, 0.7, 0.8)
    out = [
        ("union", bld.Union(s1, s2)), ("difference", bld.Difference(s1, s2)),
        ("intersection", bld.Intersection(s1, s2)), ("xor", bld.Xor(s1, s2)),
        ("smoothunion", bld.SmoothUnion(0.1, s1, s2)), ("smoothdiff", bld.SmoothDifference(0.1, s1, s2)),
        ("smoothintersect", bld.SmoothIntersect(0.1, s1, s2)),
        ("union3", bld.Union(s1, s2, bld.Translate(bld.NewSphere(0.4), -0.9, 0.2, 0.1))),
    ]
    return out


def unary3d(bld):
    s2 = bld.NewBox(1, 0.61, 0.8, 0.3)
    s2d = bld.NewRectangle(1, 0.57)
    mn, mx = s2.Bounds()
    size = mx - mn
    thickness = float(min(size.max() / 128, 0.37))
    shell = bld.Shell(s2, thickness)
    halfbox = bld.NewBox(size[0] * 20, size[1] / 3, size[2] * 20, 0)
    halfbox = bld.Translate(bld.Translate(halfbox, 0, size[1] / 3, 0), 0, size[1] / 3, 0)
    return [
        ("rotate", bld.Rotate(s2, 0.73, (1.3, 2.1, 0.4))),
        ("rotate_neg", bld.Rotate(s2, -0.31, (0.2, 0.1, 2.9))),
        ("shell", bld.Difference(shell, halfbox)),
        ("elongate", bld.Elongate(s2, 0.11, 0.27, 0.05)),
        ("round", bld.Offset(s2, -0.12)),
        ("scale", bld.Scale(s2, 1.73)),
        ("scale_small", bld.Scale(s2, 0.05)),
        ("symmetry_x", bld.Symmetry(bld.Translate(s2, 0.3, 0.2, 0.1), True, False, False)),
        ("symmetry_yz", bld.Symmetry(bld.Translate(s2, 0.3, 0.2, 0.1), False, True, True)),
        ("translate", bld.Translate(s2, 0.7, -1.1, 0.35)),
        ("array", bld.Array(s2, 1.05, 0.7, 0.93, 3, 2, 4)),
        ("circarray", bld.CircularArray(bld.Translate(s2, 1.4, 0, 0), 5, 7)),
        ("circarray_full", bld.CircularArray(bld.Translate(s2, 1.4, 0, 0), 6, 6)),
        ("twist", bld.Twist(s2, 0.61)),
        ("extrude", bld.Extrude(s2d, 1.7)),
        ("revolve", bld.Revolve(s2d, 0)),
        ("revolve_off", bld.Revolve(bld.Translate2D(s2d, 2, 0), 0.5)),
        ("transform", bld.Transform(s2, [[1, 0.2, 0, 0.3], [0, 1.1, 0.1, -0.2], [0.1, 0, 0.9, 0.5], [0, 0, 0, 1]])),
    ]


def nagon(n, r):
    """ms2.PolygonBuilder.Nagon(n, r) vertices as an (n,2) float32 array (iterated rotation, like the builder)."""
    a = np.float32(2 * math.pi) / np.float32(n)
    c, s = np.float32(math.cos(a)), np.float32(math.sin(a))
    v = np.array([r, 0], dtype=np.float32)
    out = []
    for _ in range(n):
        out.append(v.copy())
        v = np.array([c * v[0] - s * v[1], s * v[0] + c * v[1]], dtype=np.float32)
    return np.array(out, dtype=np.float32)


def primitives2d(bld):
    maxdim = 1.0
    dx, dy = maxdim, maxdim * 0.47
    thick = maxdim / 10
    verts = nagon(8, 1)
    segs = np.array([[verts[i - 1], verts[i]] for i in range(len(verts))], dtype=np.float32)
    poly = bld.NewPolygon(verts)
    return [
        ("circle", bld.NewCircle(maxdim)),
        ("line", bld.NewLine2D(0, 0, dx, dy, thick)),
        ("rect", bld.NewRectangle(dx, dy)),
        ("arc", bld.NewArc(dx, math.pi / 3, thick)),
        ("hexagon", bld.NewHexagon(maxdim)),
        ("eqtri", bld.NewEquilateralTriangle(maxdim)),
        ("poly", poly),
        ("poly_selfclosed", bld.NewPolygon([[0, 0], [0, 1], [1, 1], [0, 0]])),
        ("lines", bld.NewLines2D(segs, 0.1)),
        ("displace", bld.TranslateMulti2D(poly, verts)),
        ("octagon", bld.NewOctagon(dx)),
        ("diamond", bld.NewDiamond2D(dx, dy)),
        ("roundx", bld.NewRoundedX(dx, thick)),
        ("ellipse", bld.NewEllipse(1, 2)),
        ("ellipse_wide", bld.NewEllipse(1.7, 0.6)),
        ("bezier", bld.NewQuadraticBezier2D((dx, dy), (dx + maxdim, dy), (dx, dy + maxdim), thick)),
        ("bezier_flat", bld.NewQuadraticBezier2D((-1, 0), (0.05, 0.4), (1.2, 0.1), 0.05)),
        ("union_lines", bld.Union2D(bld.NewLine2D(1, 2, 3, 4, 0.5), bld.NewLine2D(2, 3, 0, 0, 0.2), bld.NewLine2D(2, 3, 4, 5, 0.2),
                                    bld.NewLines2D([[[0, 0], [1, 1]], [[2, 2], [3, 1]]], 0.5))),
    ]


def binops2d(bld):
    s2 = bld.NewRectangle(1, 0.61)
    s1 = bld.Translate2D(bld.NewCircle(0.4), 0.45, 1)
    return [("union2d", bld.Union2D(s1, s2)), ("diff2d", bld.Difference2D(s1, s2)),
            ("intersect2d", bld.Intersection2D(s1, s2)), ("xor2d", bld.Xor2D(s1, s2))]


def unary2d(bld):
    obj = bld.Translate2D(bld.NewRectangle(1, 0.61), 2, .3)
    return [
        ("array2d", bld.Array2D(obj, 0.83, 1.07, 3, 5)),
        ("circarray2d", bld.CircularArray2D(obj, 4, 9)),
        ("symmetry2d_x", bld.Symmetry2D(obj, True, False)),
        ("symmetry2d_xy", bld.Symmetry2D(obj, True, True)),
        ("rotate2d", bld.Rotate2D(obj, 1.234)),
        ("annulus", bld.Annulus(obj, 0.21)),
        ("offset2d", bld.Offset2D(obj, -0.13)),
        ("scale2d", bld.Scale2D(obj, 0.77)),
        ("elongate2d", bld.Elongate2D(obj, 0.4, 0.15)),
    ]


def threads3d(bld):
    T = gsdf.threads
    return [
        ("screw_iso", T.Screw(bld, 5, T.ISO(1, 0.1, True))),          # examples/test/glsdf3test.go:100-114
        ("hexhead", T.HexHead(bld, 3.4641018, 2.8867514, False, True)),
        ("nut_npt_circ", T.Nut(bld, T.NPT(0.5), T.NutCircular)),
        ("nut_iso_hex", T.Nut(bld, T.ISO(3, 0.5, False), T.NutHex)),
    ]


def guards3d(bld):
    """Screw / extrude nodes as the later operand of difference, union and smooth union -- the positions where the
    flattener plants slab guards (include/gsdf_program.h) -- directly and through translate / transform / symmetry."""
    T = gsdf.threads
    star = bld.NewPolygon(nagon(7, 0.9))
    ext = bld.Extrude(star, 0.6)
    screw = T.Screw(bld, 2.0, T.ISO(1, 0.25, True))
    return [
        ("diff_box_extrude", bld.Difference(bld.NewBox(2.4, 2.4, 2.0, 0.1), bld.Translate(ext, 0.2, 0.1, 0.5))),
        ("union_sphere_extrude", bld.Union(bld.Translate(ext, 0, 0, -1.2), bld.NewSphere(0.7), bld.Translate(ext, 0.3, 0, 1.4))),
        ("smoothunion_cyl_extrude", bld.SmoothUnion(0.25, bld.NewCylinder(0.5, 3.0, 0.05), bld.Translate(ext, 0, 0, 0.9))),
        ("smoothunion_extrude_first", bld.SmoothUnion(0.25, bld.Translate(ext, 0, 0, 0.9), bld.NewCylinder(0.5, 3.0, 0.05))),
        ("diff_cyl_rotated_screw", bld.Difference(bld.NewCylinder(1.1, 3.0, 0), bld.Rotate(screw, 0.4, (1, 0.2, 0)))),
        ("union_screw_symmetry", bld.Union(bld.NewBox(0.8, 0.8, 3.5, 0), bld.Symmetry(bld.Translate(screw, 1.3, 0, 0), True, False, False))),
        ("nested_guards", bld.Union(bld.NewSphere(0.5), bld.Translate(bld.Difference(bld.NewCylinder(0.9, 1.2, 0), screw), 0, 0, 1.6))),
        ("guard_under_scale", bld.Scale(bld.Difference(bld.NewSphere(1.2), bld.Translate(ext, 0, 0, 0.8)), 2.5)),
        ("union_two_extrudes", bld.Union(ext, bld.Translate(bld.Extrude(bld.NewCircle(0.4), 2.5), 0.2, 0, 0))),
    ]


def threads2d(bld):
    T = gsdf.threads
    return [("iso_ext_profile", T.Thread(bld, T.ISO(1, 0.1, True))), ("npt_profile", T.Thread(bld, T.NPT(0.5)))]


def scenes3d(bld):
    return [("npt-flange", gsdf.scene(bld, "npt-flange")), ("bolt", gsdf.scene(bld, "bolt")),
            ("knurled-cylinder", gsdf.scene(bld, "knurled-cylinder"))]


def geb(bld, share=True):
    """TestTransformDuplicateBug (gsdf_test.go:90-133): three extruded, non-uniformly scaled, offset letters, each used
    TWICE (a DAG, not a tree) under rotations and intersections. share=False builds the same shape from separately
    constructed, identical nodes."""
    def letter():
        s3 = bld.Extrude(bld.NewCircle(1), 1.0)
        s3 = bld.Transform(s3, [[1.2, 0, 0, 0], [0, 1.3, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])  # ms3.ScalingMat4({1.2, 1.3, 1})
        return bld.Offset(s3, -0.025)
    if share:
        G3, E3, B3 = letter(), letter(), letter()
        g, e, b = (lambda: G3), (lambda: E3), (lambda: B3)
    else:
        g = e = b = letter
    deg90 = math.pi / 2
    geb1 = bld.Intersection(bld.Intersection(g(), bld.Rotate(e(), deg90, (0, 1, 0))), bld.Rotate(b(), -deg90, (1, 0, 0)))
    geb2 = bld.Intersection(bld.Intersection(e(), bld.Rotate(g(), deg90, (0, 1, 0))), bld.Rotate(b(), -deg90, (1, 0, 0)))
    mn, mx = geb2.Bounds()
    geb2 = bld.Translate(geb2, 0, 0, float((mx - mn)[2] * np.float32(1.5)))
    return bld.Union(geb1, geb2)


def dag3d(bld):
    """Shapes whose nodes are shared between several parents."""
    s = bld.NewBox(1, 0.6, 0.8, 0.1)
    twice = bld.Union(bld.Translate(s, 1.2, 0, 0), bld.Rotate(s, 0.8, (0, 0, 1)), bld.Scale(s, 0.5))
    return [("geb_shared", geb(bld, True)), ("shared_box_union", twice),
            ("shared_smooth", bld.SmoothUnion(0.2, twice, bld.Translate(twice, 0, 0, 0.9)))]


def overlap2d(bld):
    """Unions and differences of bounded 2-D shapes that OVERLAP: the operands' bounding boxes bound their values from below
    only outside the boxes, so a box guard (include/gsdf_program.h) must never fire for a tile with points inside the
    box. Found by the random-tree fuzz on the CPU model (tests/progsim.py); the text scenes never overlap."""
    c = lambda r, x, y: bld.Translate2D(bld.NewCircle(r), x, y)
    r = lambda w, h, x, y: bld.Translate2D(bld.NewRectangle(w, h), x, y)
    star = bld.NewPolygon(nagon(5, 0.8))
    return [
        ("union_overlapping_circles", bld.Union2D(c(0.5, 0, 0), c(0.45, 0.3, 0.1), c(0.4, 0.15, 0.35), c(0.2, 1.5, 0))),
        ("union_nested", bld.Union2D(c(0.9, 0, 0), c(0.3, 0.1, 0.1), r(0.4, 0.4, -0.2, 0.2), star)),
        ("diff_line_minus_circle", bld.Difference2D(bld.NewLine2D(-0.9, -0.5, 0.3, 0.27, 0.06), c(0.74, -0.28, -0.23))),
        ("diff_small_minus_big", bld.Difference2D(c(0.3, 0.5, 0), r(1.2, 1.2, 0, 0))),
        ("diff_of_unions", bld.Difference2D(bld.Union2D(c(0.5, 0, 0), r(0.6, 0.3, 0.6, 0), star), bld.Union2D(c(0.3, 0.2, 0), c(0.25, 0.7, 0.1)))),
        ("extruded_overlap", bld.Extrude(bld.Difference2D(bld.NewRoundedX(1.0, 0.07), c(0.87, -0.13, 0.5)), 1.7)),
    ]


def overlap2d_points(name, s):
    """Evaluation points for overlap2d() shapes that put WHOLE 2048-point device tiles into the region where a box
    guard voting with inside-the-box points would skip a live operand."""
    pos = sample_points(s, dense=[128, 128] if s.is2d else [64, 64, 8])
    if name == "diff_small_minus_big":   # outside the small circle, inside the big rectangle: max(a, -s) = -s > a there
        g = append_grid(np.float32([-0.55, -0.55]), np.float32([0.55, 0.55]), [128, 128])
        g = g[np.hypot(g[:, 0] - 0.5, g[:, 1]) - 0.3 > 0.02]
        pos = np.concatenate([g[:8192], pos]).astype(np.float32)
    return pos


def all3d(bld):
    return primitives3d(bld) + binops3d(bld) + unary3d(bld) + threads3d(bld) + scenes3d(bld) + guards3d(bld)


def all2d(bld):
    return primitives2d(bld) + binops2d(bld) + unary2d(bld) + threads2d(bld)


def grid_counts(size, testres=1.0 / 3):
    """shaderTestConfig.div (gsdf_test.go:60-73): n = int(clamp(dim/testres, 5, 32)) per axis."""
    return [int(min(max(float(d) / testres, 5), 32)) for d in size]


def append_grid(mn, mx, counts):
    """ms3.AppendGrid / ms2.AppendGrid: nx*ny(*nz) points, endpoints inclusive, x fastest."""
    axes = [np.linspace(float(a), float(b), n, dtype=np.float64).astype(np.float32) for a, b, n in zip(mn, mx, counts)]
    if len(axes) == 3:
        z, y, x = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
        return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.float32)
    y, x = np.meshgrid(axes[1], axes[0], indexing="ij")
    return np.stack([x.ravel(), y.ravel()], axis=1).astype(np.float32)


def sample_points(shader, margin=0.15, dense=None):
    """The lattice testShader3D/2D evaluates (gsdf_test.go:429-436), widened by `margin` so outside points are hit too."""
    mn, mx = shader.Bounds()
    size = mx - mn
    lo, hi = mn - margin * size, mx + margin * size
    counts = dense if dense is not None else grid_counts(size)
    return append_grid(lo, hi, counts)


# ---------------------------------------------------------------------------------------------- random trees
def _rand2d(bld, rng, depth, rich=False):
    """A random 2-D tree of at most `depth` operation levels over random primitives (parameters inside the ranges the
    reference's randomised tests draw from, gsdf_test.go:572-730). rich=True (CPU fuzz only; the draws of the default
    mode are unchanged) adds line sets, multi-translations and thread profiles."""
    u = lambda a, b: float(rng.uniform(a, b))
    if rich and rng.random() < 0.15:
        k = int(rng.integers(0, 4))
        T = gsdf.threads
        if k == 0:
            pts = rng.uniform(-1, 1, (int(rng.integers(2, 6)), 2, 2)).astype(np.float32)
            return bld.NewLines2D(pts, u(0.05, 0.3))
        if k == 1:
            return bld.TranslateMulti2D(_rand2d(bld, rng, min(depth, 1), rich), rng.uniform(-1.5, 1.5, (int(rng.integers(1, 5)), 2)).astype(np.float32))
        if k == 2:
            if rng.random() < 0.5:
                return bld.NewEllipse(u(0.3, 1.5), u(0.3, 1.5)) if rng.random() < 0.5 else \
                    bld.NewQuadraticBezier2D((u(-1, 0), u(-1, 0)), (u(-0.5, 0.5), u(0.2, 1)), (u(0.2, 1), u(-0.5, 0.5)), u(0.03, 0.2))
            return T.Thread(bld, T.ISO(u(0.8, 3.0), u(0.1, 0.5), bool(rng.integers(0, 2))))
        return T.Thread(bld, [T.NPT(0.5), T.UTS(0.5, 13, True), T.Acme(1.0, 0.2), T.PlasticButtress(1.0, 0.2)][int(rng.integers(0, 4))])
    if depth <= 0 or rng.random() < 0.1:
        k = int(rng.integers(0, 10))
        if k == 0: return bld.NewCircle(u(0.3, 1.2))
        if k == 1: return bld.NewRectangle(u(0.3, 1.5), u(0.3, 1.5))
        if k == 2: return bld.NewHexagon(u(0.3, 1.0))
        if k == 3: return bld.NewOctagon(u(0.3, 1.0))
        if k == 4: return bld.NewEquilateralTriangle(u(0.4, 1.2))
        if k == 5: return bld.NewPolygon(nagon(int(rng.integers(3, 12)), np.float32(u(0.4, 1.2))))
        if k == 6: return bld.NewLine2D(u(-1, 0), u(-1, 0), u(0.1, 1), u(0.1, 1), u(0.05, 0.3))
        if k == 7: return bld.NewDiamond2D(u(0.3, 1.2), u(0.3, 1.2))
        if k == 8: return bld.NewArc(u(0.5, 1.2), u(0.3, 2.5), u(0.05, 0.2))
        return bld.NewRoundedX(u(0.5, 1.2), u(0.05, 0.2))
    k = int(rng.integers(0, 13))
    a = _rand2d(bld, rng, depth - 1, rich)
    if k == 0: return bld.Union2D(a, _rand2d(bld, rng, depth - 1, rich), *[_rand2d(bld, rng, 0, rich) for _ in range(int(rng.integers(0, 3)))])
    if k == 1: return bld.Difference2D(a, bld.Translate2D(_rand2d(bld, rng, depth - 1, rich), u(-0.5, 0.5), u(-0.5, 0.5)))
    if k == 2: return bld.Intersection2D(a, _rand2d(bld, rng, depth - 1, rich))
    if k == 3: return bld.Xor2D(a, bld.Translate2D(_rand2d(bld, rng, depth - 1, rich), u(-0.5, 0.5), u(-0.5, 0.5)))
    if k == 4: return bld.Translate2D(a, u(-1.5, 1.5), u(-1.5, 1.5))
    if k == 5: return bld.Rotate2D(a, u(-3, 3))
    if k == 6: return bld.Scale2D(a, u(0.3, 2.5))
    if k == 7: return bld.Offset2D(a, u(-0.1, 0.2))
    if k == 8: return bld.Annulus(a, u(0.02, 0.15))
    if k == 9: return bld.Symmetry2D(bld.Translate2D(a, u(0, 1), u(0, 1)), bool(rng.integers(0, 2)), True)
    if k == 10: return bld.Elongate2D(a, u(0.05, 0.6), u(0.05, 0.6))
    if k == 11: return bld.Array2D(a, u(1.5, 3.0), u(1.5, 3.0), int(rng.integers(1, 4)), int(rng.integers(1, 4)))
    div = int(rng.integers(3, 10))
    return bld.CircularArray2D(bld.Translate2D(a, u(1.5, 3.0), 0), int(rng.integers(1, div + 1)), div)


def _rand3d(bld, rng, depth, rich=False):
    u = lambda a, b: float(rng.uniform(a, b))
    if rich and rng.random() < 0.12:
        k = int(rng.integers(0, 5))
        T = gsdf.threads
        if k == 0: return T.Nut(bld, T.ISO(u(2, 4), 0.5, False), [T.NutHex, T.NutCircular][int(rng.integers(0, 2))])
        if k == 1: return T.HexHead(bld, u(1, 3), u(1, 3), bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))
        if k == 2: return T.Screw(bld, u(1.0, 3.0), [T.NPT(0.5), T.UTS(0.5, 13, True), T.Acme(1.0, 0.2), T.ANSIButtress(1.0, 0.2)][int(rng.integers(0, 4))])
        if k == 3:
            inner = _rand3d(bld, rng, min(depth, 1), rich)
            mn, mx = inner.Bounds()
            return bld.OverloadShader3DBounds(inner, mn - 0.1, mx + 0.2)
        return bld.NewBoundsBoxFrame(np.float32([-u(0.5, 1), -u(0.5, 1), -u(0.5, 1)]), np.float32([u(0.5, 1), u(0.5, 1), u(0.5, 1)]))
    if depth <= 0 or rng.random() < 0.2:
        k = int(rng.integers(0, 10))
        if k == 0: return bld.NewSphere(u(0.3, 1.2))
        if k == 1: return bld.NewBox(u(0.5, 1.5), u(0.5, 1.5), u(0.5, 1.5), u(0, 0.12))
        if k == 2: return bld.NewCylinder(u(0.3, 1.0), u(0.5, 2.0), 0 if rng.random() < 0.5 else u(0.01, 0.1))
        if k == 3: return bld.NewHexagonalPrism(u(0.4, 1.2), u(0.4, 1.5))
        if k == 4: return bld.NewTorus(u(0.8, 1.5), u(0.1, 0.35))
        if k == 5: return bld.NewBoxFrame(u(0.8, 1.5), u(0.8, 1.5), u(0.8, 1.5), u(0.05, 0.15))
        if k == 6: return bld.NewTriangularPrism(u(0.5, 1.2), u(0.3, 1.5))
        if k == 7: return bld.Extrude(_rand2d(bld, rng, 1, rich), u(0.3, 2.0))
        if k == 8: return bld.Revolve(bld.Translate2D(_rand2d(bld, rng, 1, rich), u(1.5, 3.0), 0), 0 if rng.random() < 0.5 else u(0.1, 0.5))
        T = gsdf.threads
        return T.Screw(bld, u(1.0, 3.0), T.ISO(u(0.8, 1.6), u(0.1, 0.3), bool(rng.integers(0, 2))))
    k = int(rng.integers(0, 18))
    a = _rand3d(bld, rng, depth - 1, rich)
    other = lambda: bld.Translate(_rand3d(bld, rng, depth - 1, rich), u(-0.6, 0.6), u(-0.6, 0.6), u(-0.6, 0.6))
    if k == 0: return bld.Union(a, other(), *[_rand3d(bld, rng, 0, rich) for _ in range(int(rng.integers(0, 3)))])
    if k == 1: return bld.Difference(a, other())
    if k == 2: return bld.Intersection(a, other())
    if k == 3: return bld.Xor(a, other())
    if k == 4: return bld.SmoothUnion(u(0.05, 0.4), a, other())
    if k == 5: return bld.SmoothDifference(u(0.05, 0.4), a, other())
    if k == 6: return bld.SmoothIntersect(u(0.05, 0.4), a, other())
    if k == 7: return bld.Translate(a, u(-1.5, 1.5), u(-1.5, 1.5), u(-1.5, 1.5))
    if k == 8: return bld.Rotate(a, u(-3, 3), (u(0.1, 1), u(-1, 1), u(-1, 1)))
    if k == 9: return bld.Scale(a, u(0.2, 3.0))
    if k == 10: return bld.Offset(a, u(-0.05, 0.15))
    if k == 11: return bld.Shell(a, u(0.02, 0.1))
    if k == 12: return bld.Elongate(a, u(0.05, 0.5), u(0.05, 0.5), u(0.05, 0.5))
    if k == 13: return bld.Symmetry(bld.Translate(a, u(0, 1), u(0, 1), u(0, 1)), True, bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))
    if k == 14: return bld.Twist(a, u(-0.8, 0.8))
    if k == 15: return bld.Array(a, u(2.0, 4.0), u(2.0, 4.0), u(2.0, 4.0), int(rng.integers(1, 3)), int(rng.integers(1, 3)), int(rng.integers(1, 3)))
    if k == 16:
        div = int(rng.integers(3, 9))
        return bld.CircularArray(bld.Translate(a, u(1.5, 3.0), 0, 0), int(rng.integers(1, div + 1)), div)
    return bld.Transform(a, [[u(0.8, 1.2), u(-0.2, 0.2), 0, u(-0.5, 0.5)], [0, u(0.8, 1.2), u(-0.2, 0.2), u(-0.5, 0.5)],
                             [u(-0.2, 0.2), 0, u(0.8, 1.2), u(-0.5, 0.5)], [0, 0, 0, 1]])


def random_trees(bld, seed, count, dim=3, depth=4, max_dstack=16, max_pstack=8, rich=False):
    """`count` seeded random trees (numpy Generator, PCG64: reproducible everywhere) that fit the interpreter's stacks."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        s = (_rand3d if dim == 3 else _rand2d)(bld, rng, depth, rich)
        f = bld.flatten(s)
        if f["dstack"] > max_dstack or f["pstack"] > max_pstack:
            continue
        out.append(("rand%dd_%d_%d" % (dim, seed, len(out)), s))
    return out
