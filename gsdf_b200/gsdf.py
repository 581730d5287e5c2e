"""Mirror of the reference's root package `gsdf` (gsdf.Builder, primitives*.go, operations*.go) and of
`forge/threads`, backed by the C++ host layer in libgsdfb200.so.

Method names and argument order are the reference's (Builder.NewSphere, Builder.Union, Builder.Translate, ...), so
tests read like gsdf_test.go. A `Shader` is a node of the CSG tree (glbuild.Shader3D / Shader2D): it has Bounds()
and is handed to gleval.NewCUDASDF3 / glrender to be evaluated on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib

# gsdf_node_kind (include/gsdf_tree.h)
K = dict(
    SPHERE=1, BOX=2, CYLINDER=3, HEX=4, TORUS=5, BOXFRAME=6,
    UNION=16, DIFF=17, INTERSECT=18, XOR=19, SMOOTH_UNION=20, SMOOTH_DIFF=21, SMOOTH_INTERSECT=22, SCALE=23,
    SYMMETRY=24, TRANSFORM=25, TRANSLATE=26, OFFSET=27, ARRAY=28, ELONGATE=29, SHELL=30, CIRCARRAY=31, TWIST=32, BOUNDS3=33,
    EXTRUDE=40, REVOLVE=41, SCREW=42,
    LINE2D=64, LINES2D=65, ARC2D=66, CIRCLE2D=67, EQTRI2D=68, RECT2D=69, HEX2D=70, OCT2D=71, ELLIPSE2D=72, POLY2D=73,
    DIAMOND2D=74, ROUNDX2D=75, BEZIERQ2D=76,
    UNION2D=96, DIFF2D=97, INTERSECT2D=98, XOR2D=99, ARRAY2D=100, OFFSET2D=101, TRANSLATE2D=102, ROTATE2D=103,
    SYMMETRY2D=104, ANNULUS2D=105, CIRCARRAY2D=106, SCALE2D=107, TRANSLATEMULTI2D=108, ELONGATE2D=109, BOUNDS2=110,
    CALL_ROTATE=200, CALL_TRANSFORM16=201, CALL_TRIPRISM=202, CALL_BOUNDSBOXFRAME=203,
)

NutCircular, NutHex, NutKnurl = 1, 2, 3  # forge/threads/nut.go:12-17


class ShapeError(ValueError):
    """Raised where the reference's Builder panics (gsdf.go:100-103) unless FlagNoDimensionPanic is set."""


class Shader:
    """A node of the tree: glbuild.Shader3D or glbuild.Shader2D (glbuild/glbuild.go:63-82)."""

    def __init__(self, bld, node_id):
        self.bld = bld
        self.id = int(node_id)

    @property
    def is2d(self):
        return bool(lib.gsdfh_is2d(self.bld._h, self.id))

    def Bounds(self):
        """Shader3D.Bounds() -> (min[3], max[3]) or Shader2D.Bounds() -> (min[2], max[2]) as float32 arrays."""
        if self.is2d:
            out = (C.c_float * 4)()
            if lib.gsdfh_bounds2(self.bld._h, self.id, out) != 0:
                raise ShapeError(self.bld.Err())
            a = np.array(out, dtype=np.float32)
            return a[:2].copy(), a[2:].copy()
        out = (C.c_float * 6)()
        if lib.gsdfh_bounds3(self.bld._h, self.id, out) != 0:
            raise ShapeError(self.bld.Err())
        a = np.array(out, dtype=np.float32)
        return a[:3].copy(), a[3:].copy()

    def Diagonal(self):
        """ms3.Box.Diagonal(): Norm(Size()) in float32 (nested Hypot), as flange.go:77 uses for -resdiv."""
        mn, mx = self.Bounds()
        s = (mx - mn).astype(np.float32)
        h = _hypot32(s[1], s[2]) if len(s) == 3 else s[1]
        return _hypot32(s[0], h)


def _hypot32(p, q):
    p, q = np.float32(abs(p)), np.float32(abs(q))
    if p < q:
        p, q = q, p
    if p == 0:
        return np.float32(0)
    q = np.float32(q / p)
    return np.float32(p * np.sqrt(np.float32(np.float32(1) + np.float32(q * q))))


class Builder:
    """gsdf.Builder (gsdf.go:44). panic_on_error=True mimics the default flags (shape errors panic);
    False mimics FlagNoDimensionPanic: errors accumulate and are read with Err()."""

    def __init__(self, panic_on_error=True):
        self._h = lib.gsdfh_builder_new()
        self.panic_on_error = panic_on_error

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.gsdfh_builder_free(h)

    def Err(self):
        return lib.gsdfh_builder_err(self._h).decode()

    def ClearErrors(self):
        lib.gsdfh_builder_clear_errors(self._h)

    # ------------------------------------------------------------------ plumbing
    def _wrap(self, node_id):
        if node_id < 0 or (self.panic_on_error and self.Err()):
            msg = self.Err() or "shape construction failed"
            self.ClearErrors()
            raise ShapeError(msg)
        return Shader(self, node_id)

    def _node(self, kind, f=(), ip=(), children=(), aux=()):
        fa = (C.c_float * max(len(f), 1))(*[float(v) for v in f])
        ia = (C.c_int32 * max(len(ip), 1))(*[int(v) for v in ip])
        ids = []
        for c in children:
            if not isinstance(c, Shader) or c.bld is not self:
                raise ShapeError("nil SDF argument")  # gsdf.go:108-110
            ids.append(c.id)
        ca = (C.c_int32 * max(len(ids), 1))(*ids)
        auxf = np.ascontiguousarray(np.asarray(aux, dtype=np.float32).reshape(-1))
        ap = auxf.ctypes.data_as(C.POINTER(C.c_float))
        return self._wrap(lib.gsdfh_node(self._h, K[kind], fa, len(f), ia, len(ip), ca, len(ids), ap, auxf.size))

    # ------------------------------------------------------------------ 3D primitives (primitives.go)
    def NewSphere(self, r): return self._node("SPHERE", [r])
    def NewBox(self, x, y, z, round): return self._node("BOX", [x, y, z, round])
    def NewCylinder(self, r, h, rounding): return self._node("CYLINDER", [r, h, rounding])
    def NewHexagonalPrism(self, face2Face, h): return self._node("HEX", [face2Face, h])
    def NewTriangularPrism(self, triHeight, extrudeLength): return self._node("CALL_TRIPRISM", [triHeight, extrudeLength])
    def NewTorus(self, greaterRadius, lesserRadius): return self._node("TORUS", [greaterRadius, lesserRadius])
    def NewBoxFrame(self, dimX, dimY, dimZ, e): return self._node("BOXFRAME", [dimX, dimY, dimZ, e])
    def NewBoundsBoxFrame(self, bbmin, bbmax): return self._node("CALL_BOUNDSBOXFRAME", list(bbmin) + list(bbmax))

    # ------------------------------------------------------------------ 3D operations (operations.go)
    def Union(self, *shaders): return self._node("UNION", children=shaders)
    def Difference(self, a, b): return self._node("DIFF", children=[a, b])
    def Intersection(self, a, b): return self._node("INTERSECT", children=[a, b])
    def Xor(self, a, b): return self._node("XOR", children=[a, b])
    def Scale(self, s, scaleFactor): return self._node("SCALE", [scaleFactor], children=[s])
    def Symmetry(self, s, mirrorX, mirrorY, mirrorZ):
        return self._node("SYMMETRY", ip=[int(bool(mirrorX)) | int(bool(mirrorY)) << 1 | int(bool(mirrorZ)) << 2], children=[s])
    def Transform(self, s, m4x4): return self._node("CALL_TRANSFORM16", list(np.asarray(m4x4, dtype=np.float32).reshape(16)), children=[s])
    def Rotate(self, s, radians, axis): return self._node("CALL_ROTATE", [radians, axis[0], axis[1], axis[2]], children=[s])
    def Translate(self, s, dirX, dirY, dirZ): return self._node("TRANSLATE", [dirX, dirY, dirZ], children=[s])
    def Offset(self, s, sdfAdd): return self._node("OFFSET", [sdfAdd], children=[s])
    def Array(self, s, spacingX, spacingY, spacingZ, nx, ny, nz):
        return self._node("ARRAY", [spacingX, spacingY, spacingZ], [nx, ny, nz], [s])
    def SmoothUnion(self, k, s1, s2): return self._node("SMOOTH_UNION", [k], children=[s1, s2])
    def SmoothDifference(self, k, s1, s2): return self._node("SMOOTH_DIFF", [k], children=[s1, s2])
    def SmoothIntersect(self, k, s1, s2): return self._node("SMOOTH_INTERSECT", [k], children=[s1, s2])
    def Elongate(self, s, dirX, dirY, dirZ): return self._node("ELONGATE", [dirX, dirY, dirZ], children=[s])
    def Shell(self, s, thickness): return self._node("SHELL", [thickness], children=[s])
    def CircularArray(self, s, numInstances, circleDiv): return self._node("CIRCARRAY", ip=[numInstances, circleDiv], children=[s])
    def Twist(self, s, k): return self._node("TWIST", [k], children=[s])

    # ------------------------------------------------------------------ glbuild wrappers (glbuild/glbuild.go:1080-1128)
    def OverloadShader3DBounds(self, s, bbmin, bbmax): return self._node("BOUNDS3", list(bbmin) + list(bbmax), children=[s])
    def OverloadShader2DBounds(self, s, bbmin, bbmax): return self._node("BOUNDS2", list(bbmin) + list(bbmax), children=[s])

    # ------------------------------------------------------------------ 2D -> 3D (operations2d.go)
    def Extrude(self, s, h): return self._node("EXTRUDE", [h], children=[s])
    def Revolve(self, s, axisOffset): return self._node("REVOLVE", [axisOffset], children=[s])

    # ------------------------------------------------------------------ 2D primitives (primitives2d.go)
    def NewLine2D(self, x0, y0, x1, y1, width): return self._node("LINE2D", [x0, y0, x1, y1, width])
    def NewLines2D(self, segments, width): return self._node("LINES2D", [width], aux=np.asarray(segments, dtype=np.float32))
    def NewArc(self, radius, arcAngle, thick): return self._node("ARC2D", [radius, arcAngle, thick])
    def NewCircle(self, radius): return self._node("CIRCLE2D", [radius])
    def NewEquilateralTriangle(self, triangleHeight): return self._node("EQTRI2D", [triangleHeight])
    def NewRectangle(self, x, y): return self._node("RECT2D", [x, y])
    def NewHexagon(self, side): return self._node("HEX2D", [side])
    def NewOctagon(self, constrain): return self._node("OCT2D", [constrain])
    def NewPolygon(self, vertices): return self._node("POLY2D", aux=np.asarray(vertices, dtype=np.float32))
    def NewDiamond2D(self, x_width, y_height): return self._node("DIAMOND2D", [x_width, y_height])
    def NewRoundedX(self, width, thick): return self._node("ROUNDX2D", [width, thick])
    def NewEllipse(self, a, b): return self._node("ELLIPSE2D", [a, b])
    def NewQuadraticBezier2D(self, a, b, c, thick): return self._node("BEZIERQ2D", [a[0], a[1], b[0], b[1], c[0], c[1], thick])

    # ------------------------------------------------------------------ 2D operations (operations2d.go)
    def Union2D(self, *shaders): return self._node("UNION2D", children=shaders)
    def Difference2D(self, a, b): return self._node("DIFF2D", children=[a, b])
    def Intersection2D(self, a, b): return self._node("INTERSECT2D", children=[a, b])
    def Xor2D(self, a, b): return self._node("XOR2D", children=[a, b])
    def Array2D(self, s, spacingX, spacingY, nx, ny): return self._node("ARRAY2D", [spacingX, spacingY], [nx, ny], [s])
    def Offset2D(self, s, sdfAdd): return self._node("OFFSET2D", [sdfAdd], children=[s])
    def Translate2D(self, s, dirX, dirY): return self._node("TRANSLATE2D", [dirX, dirY], children=[s])
    def Rotate2D(self, s, theta): return self._node("ROTATE2D", [theta], children=[s])
    def Symmetry2D(self, s, mirrorX, mirrorY):
        return self._node("SYMMETRY2D", ip=[int(bool(mirrorX)) | int(bool(mirrorY)) << 1], children=[s])
    def Annulus(self, s, sub): return self._node("ANNULUS2D", [sub], children=[s])
    def CircularArray2D(self, s, numInstances, circleDiv): return self._node("CIRCARRAY2D", ip=[numInstances, circleDiv], children=[s])
    def Scale2D(self, s, scale): return self._node("SCALE2D", [scale], children=[s])
    def TranslateMulti2D(self, s, displacements):
        return self._node("TRANSLATEMULTI2D", children=[s], aux=np.asarray(displacements, dtype=np.float32))
    def Elongate2D(self, s, dirX, dirY): return self._node("ELONGATE2D", [dirX, dirY], children=[s])

    # ------------------------------------------------------------------ tree table (for the CPU oracle in tests)
    def tree_table(self):
        """Returns (nodes_bytes, children int32[], aux float32[]): copies of the gsdf_tree_node table."""
        nodes = C.POINTER(_lib.TreeNode)()
        ch = C.POINTER(C.c_int32)()
        aux = C.POINTER(C.c_float)()
        nn, nc, na = C.c_int32(), C.c_int32(), C.c_int32()
        lib.gsdfh_tree(self._h, C.byref(nodes), C.byref(nn), C.byref(ch), C.byref(nc), C.byref(aux), C.byref(na))
        nb = C.string_at(nodes, nn.value * C.sizeof(_lib.TreeNode)) if nn.value else b""
        chv = np.ctypeslib.as_array(ch, shape=(nc.value,)).copy() if nc.value else np.zeros(0, np.int32)
        auxv = np.ctypeslib.as_array(aux, shape=(na.value,)).copy() if na.value else np.zeros(0, np.float32)
        return nb, chv, auxv

    def tree_nodes(self):
        """The node table as a list of TreeNode records (kind, nchild, child_off, aux_off, aux_cnt, iparam, fparam)."""
        nb, _, _ = self.tree_table()
        n = len(nb) // C.sizeof(_lib.TreeNode)
        return list((_lib.TreeNode * n).from_buffer_copy(nb)) if n else []

    def tree_children(self):
        return self.tree_table()[1]

    def flatten(self, root):
        """Flattener output for inspection: dict(blob, aux, dim, ninstr, nchunks, dstack, pstack)."""
        f = lib.gsdfh_flatten(self._h, root.id)
        if not f:
            msg = self.Err()
            self.ClearErrors()
            raise ShapeError(msg)
        try:
            nb, na = C.c_size_t(), C.c_size_t()
            bp = lib.gsdfh_flat_blob(f, C.byref(nb))
            ap = lib.gsdfh_flat_aux(f, C.byref(na))
            info = (C.c_int32 * 5)()
            lib.gsdfh_flat_info(f, info)
            return dict(blob=C.string_at(bp, nb.value),
                        aux=np.ctypeslib.as_array(ap, shape=(na.value,)).copy() if na.value else np.zeros(0, np.float32),
                        dim=info[0], ninstr=info[1], nchunks=info[2], dstack=info[3], pstack=info[4])
        finally:
            lib.gsdfh_flat_free(f)


# ---------------------------------------------------------------------- forge/threads
class ISO:
    """threads.ISO{D, P, Ext} (forge/threads/iso.go:20-29)."""
    kind = 0

    def __init__(self, D, P, Ext=False):
        self.p0, self.p1, self.ext = float(D), float(P), bool(Ext)


class NPT:
    """threads.NPT set by SetFromNominal(nominal) (forge/threads/npt.go:63-74)."""
    kind = 1

    def __init__(self, nominal):
        self.p0, self.p1, self.ext = float(nominal), 0.0, False


class UTS:
    """threads.UTS{D, TPI, Ext} (forge/threads/uts.go:8-15)."""
    kind = 2

    def __init__(self, D, TPI, Ext=False):
        self.p0, self.p1, self.ext = float(D), float(TPI), bool(Ext)


class Acme:
    """threads.Acme{D, P} (forge/threads/acme.go:10-15)."""
    kind = 3

    def __init__(self, D, P):
        self.p0, self.p1, self.ext = float(D), float(P), False


class ANSIButtress(Acme):
    """threads.ANSIButtress{D, P} (forge/threads/ansibuttress.go:10-15)."""
    kind = 4


class PlasticButtress(Acme):
    """threads.PlasticButtress{D, P} (forge/threads/plasticbuttress.go:9-14)."""
    kind = 5


class threads:
    """Namespace mirroring package forge/threads."""
    ISO, NPT, UTS, Acme, ANSIButtress, PlasticButtress = ISO, NPT, UTS, Acme, ANSIButtress, PlasticButtress
    NutCircular, NutHex, NutKnurl = NutCircular, NutHex, NutKnurl

    @staticmethod
    def Thread(bld, t): return bld._wrap(lib.gsdfh_thread_profile(bld._h, t.kind, t.p0, t.p1, int(t.ext)))
    @staticmethod
    def Screw(bld, length, t): return bld._wrap(lib.gsdfh_screw(bld._h, length, t.kind, t.p0, t.p1, int(t.ext)))
    @staticmethod
    def Nut(bld, Thread, Style, Tolerance=0.0):
        return bld._wrap(lib.gsdfh_nut(bld._h, Thread.kind, Thread.p0, Thread.p1, int(Thread.ext), Style, Tolerance))
    @staticmethod
    def Bolt(bld, Thread, Style, TotalLength, ShankLength, Tolerance=0.0):
        return bld._wrap(lib.gsdfh_bolt(bld._h, Thread.kind, Thread.p0, Thread.p1, int(Thread.ext), Style, Tolerance, TotalLength, ShankLength))
    @staticmethod
    def HexHead(bld, radius, height, roundNeg, roundPos):
        return bld._wrap(lib.gsdfh_hexhead(bld._h, radius, height, int(roundNeg), int(roundPos)))


def scene(bld, name, param=0.0):
    """The example programs' scene() functions: 'npt-flange' (examples/npt-flange/flange.go:23), 'bolt'
    (examples/bolt/main.go:26), 'knurled-cylinder' (examples/knurled-cylinder/knurled-cyl.go:57, param = -d),
    'fibonacci-showerhead' (examples/fibonacci-showerhead/showerhead.go:31)."""
    return bld._wrap(lib.gsdfh_scene(bld._h, name.encode(), float(param)))
