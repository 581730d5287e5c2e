"""ctypes binding of the CPU ORACLE (oracle/gsdf_oracle.c). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product (gsdf_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libgsdf_oracle.so")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("gsdf_oracle.c", "gsdf_oracle_dc.c", "gsdf_oracle.h", "mc_tables.inc", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return LIB_PATH


class GoNode(C.Structure):  # go_node, 96 bytes (same layout as gsdf_tree_node)
    _fields_ = [("kind", C.c_int32), ("nchild", C.c_int32), ("child_off", C.c_int32), ("aux_off", C.c_int32),
                ("aux_cnt", C.c_int32), ("iparam", C.c_int32 * 3), ("fparam", C.c_float * 16)]


class GoTree(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("nnodes", C.c_int32), ("children", C.c_void_p), ("aux", C.c_void_p), ("root", C.c_int32)]


class GoLattice(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("res", C.c_float), ("n", C.c_int32 * 3)]


class GoColorConv(C.Structure):  # go_colorconv: same layout as gsdf_colorconv
    _fields_ = [("kind", C.c_int32), ("p", C.c_float * 7), ("c0", C.c_uint32), ("c1", C.c_uint32)]


def colorconv(kind, p=(), c0=0, c1=0):
    cc = GoColorConv()
    cc.kind = kind
    for i, v in enumerate(p):
        cc.p[i] = float(v)
    cc.c0, cc.c1 = int(c0), int(c1)
    return cc


def colorconv_linear_gradient(length, rgba0, rgba1):
    """gsdfaux.ColorConversionLinearGradient (color.go:51-73); colours as R|G<<8|B<<16|A<<24."""
    if rgba0 == 0xff000000 and rgba1 == 0xffffffff:
        return colorconv(1, [length])
    hsv = []
    for c in (rgba0, rgba1):
        out = (C.c_float * 3)()
        f = np.float32
        lib().go_rgb_to_hsv(f(c & 255) / f(255), f((c >> 8) & 255) / f(255), f((c >> 16) & 255) / f(255), out)
        hsv += list(out)
    return colorconv(3, hsv + [length], rgba0, rgba1)


def colorconv_inigo_quilez(characteristic):
    return colorconv(2, [np.float32(1) / np.float32(characteristic)])


def color_of(cc, d):
    return np.array([lib().go_color_of(C.byref(cc), float(v)) for v in np.asarray(d, np.float32).reshape(-1)], np.uint32)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        f32p, vp = C.POINTER(C.c_float), C.c_void_p
        for name in ("go_sqrt", "go_atan", "go_sin", "go_cos", "go_tan", "go_floor", "go_round", "go_acos", "go_cbrt", "go_log", "go_exp"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_float]
        for name in ("go_hypot", "go_atan2", "go_min", "go_max"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_float, C.c_float]
        L.go_eval3.restype = C.c_int
        L.go_eval3.argtypes = [C.POINTER(GoTree), vp, vp, C.c_size_t]
        L.go_eval2.restype = C.c_int
        L.go_eval2.argtypes = [C.POINTER(GoTree), vp, vp, C.c_size_t]
        L.go_flat_lattice.restype = C.c_int
        L.go_flat_lattice.argtypes = [f32p, f32p, C.c_float, C.POINTER(GoLattice)]
        L.go_octree_levels.restype = C.c_int
        L.go_octree_levels.argtypes = [f32p, f32p, C.c_float]
        L.go_flat_eval_grid.restype = C.c_int64
        L.go_flat_eval_grid.argtypes = [C.POINTER(GoTree), C.POINTER(GoLattice), vp, C.c_int, C.c_int]
        L.go_flat_eval_planes.restype = C.c_int64
        L.go_flat_eval_planes.argtypes = [C.POINTER(GoTree), C.POINTER(GoLattice), C.c_int, C.c_int, vp, C.c_int, C.c_int]
        L.go_flat_march.restype = C.c_int64
        L.go_flat_march.argtypes = [C.POINTER(GoLattice), vp, vp, C.c_int64, vp, vp]
        L.go_flat_march_slab.restype = C.c_int64
        L.go_flat_march_slab.argtypes = [C.POINTER(GoLattice), vp, vp, C.c_int64, vp, vp, C.c_int, C.c_int]
        L.go_flat_march_slab_w.restype = C.c_int64
        L.go_flat_march_slab_w.argtypes = [C.POINTER(GoLattice), vp, vp, C.c_int64, vp, vp, C.c_int, C.c_int, C.c_int]
        L.go_octree_prune_mask.restype = C.c_int64
        L.go_octree_prune_mask.argtypes = [C.POINTER(GoTree), C.POINTER(GoLattice), vp]
        L.go_octree_prune_plan.restype = C.c_int64
        L.go_octree_prune_plan.argtypes = [C.POINTER(GoTree), C.POINTER(GoLattice), C.c_int, C.POINTER(C.c_int), f32p, vp, C.POINTER(C.c_int64)]
        L.go_mc_cube.restype = C.c_int
        L.go_mc_cube.argtypes = [f32p, f32p, f32p, C.POINTER(C.c_int)]
        L.go_stl_write.restype = C.c_int64
        L.go_stl_write.argtypes = [vp, C.c_int64, vp]
        L.go_stl_read.restype = C.c_int64
        L.go_stl_read.argtypes = [vp, C.c_size_t, vp, C.c_int64]
        L.go_image_eval2.restype = C.c_int
        L.go_image_eval2.argtypes = [C.POINTER(GoTree), f32p, f32p, C.c_int, C.c_int, vp]
        L.go_color_of.restype = C.c_uint32
        L.go_color_of.argtypes = [C.POINTER(GoColorConv), C.c_float]
        L.go_rgb_to_hsv.restype = None
        L.go_rgb_to_hsv.argtypes = [C.c_float, C.c_float, C.c_float, f32p]
        L.go_image_render2.restype = C.c_int
        L.go_image_render2.argtypes = [C.POINTER(GoTree), f32p, f32p, C.c_int, C.c_int, C.POINTER(GoColorConv), vp]
        L.go_dc_levels.restype = C.c_int
        L.go_dc_levels.argtypes = [f32p, f32p, C.c_float, f32p]
        L.go_dual_contour.restype = C.c_int64
        L.go_dual_contour.argtypes = [C.POINTER(GoTree), f32p, f32p, C.c_float, C.c_int, vp, C.c_int64, C.POINTER(C.c_int64)]
        L.go_dual_contour_ex.restype = C.c_int64
        L.go_dual_contour_ex.argtypes = [C.POINTER(GoTree), f32p, f32p, C.c_float, C.c_int, vp, C.c_int64, C.POINTER(C.c_int64), vp]
        L.go_lsq_mgs64.restype = C.c_int
        L.go_lsq_mgs64.argtypes = [C.c_int, f32p, f32p, f32p]
        L.go_mc_edge_table.restype = C.POINTER(C.c_int)
        L.go_mc_tri_table.restype = C.POINTER(C.c_int8)
        L.go_mc_pair_table.restype = C.POINTER(C.c_int)
        _lib = L
    return _lib


class Tree:
    """A tree table (bytes of go_node records, children int32[], aux float32[], root id) kept alive for the C calls."""

    def __init__(self, nodes_bytes, children, aux, root):
        self._nodes = C.create_string_buffer(nodes_bytes, len(nodes_bytes))
        self._children = np.ascontiguousarray(children, dtype=np.int32)
        self._aux = np.ascontiguousarray(aux, dtype=np.float32)
        if self._children.size == 0:
            self._children = np.zeros(1, np.int32)
        if self._aux.size == 0:
            self._aux = np.zeros(1, np.float32)
        self.c = GoTree(C.cast(self._nodes, C.c_void_p), len(nodes_bytes) // C.sizeof(GoNode),
                        self._children.ctypes.data, self._aux.ctypes.data, int(root))

    @classmethod
    def from_shader(cls, shader):
        """Builds the oracle's tree table from a gsdf_b200.gsdf Shader (the host-side tree, NOT the flattened program)."""
        nb, ch, aux = shader.bld.tree_table()
        return cls(nb, ch, aux, shader.id)

    def eval3(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        dist = np.empty(len(pos), dtype=np.float32)
        rc = lib().go_eval3(C.byref(self.c), pos.ctypes.data, dist.ctypes.data, len(pos))
        if rc:
            raise RuntimeError("oracle go_eval3 failed: %d" % rc)
        return dist

    def eval2(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 2)
        dist = np.empty(len(pos), dtype=np.float32)
        rc = lib().go_eval2(C.byref(self.c), pos.ctypes.data, dist.ctypes.data, len(pos))
        if rc:
            raise RuntimeError("oracle go_eval2 failed: %d" % rc)
        return dist

    def image_eval2(self, bbmin, bbmax, w, h):
        out = np.empty((h, w), dtype=np.float32)
        a = (C.c_float * 2)(float(bbmin[0]), float(bbmin[1]))
        b = (C.c_float * 2)(float(bbmax[0]), float(bbmax[1]))
        rc = lib().go_image_eval2(C.byref(self.c), a, b, w, h, out.ctypes.data)
        if rc:
            raise RuntimeError("oracle go_image_eval2 failed: %d" % rc)
        return out


def image_render2(tree, bbmin, bbmax, w, h, cc=None):
    """ImageRendererSDF2.Render: uint8 (h, w, 4)."""
    out = np.empty((h, w, 4), dtype=np.uint8)
    a = (C.c_float * 2)(float(bbmin[0]), float(bbmin[1]))
    b = (C.c_float * 2)(float(bbmax[0]), float(bbmax[1]))
    rc = lib().go_image_render2(C.byref(tree.c), a, b, w, h, C.byref(cc) if cc is not None else None, out.ctypes.data)
    if rc:
        raise RuntimeError("oracle go_image_render2 failed: %d" % rc)
    return out


DC_NAIVE, DC_LSQ, DC_LSQ_CHISELED = 0, 1, 2


def dual_contour(tree, bbmin, bbmax, res, placer=DC_LSQ):
    """DualContourRenderer.Reset + RenderAll: returns (triangles (n,3,3), stats dict)."""
    a = (C.c_float * 3)(*[float(v) for v in bbmin])
    b = (C.c_float * 3)(*[float(v) for v in bbmax])
    st = (C.c_int64 * 4)()
    n = lib().go_dual_contour(C.byref(tree.c), a, b, float(res), placer, None, 0, st)
    if n < 0:
        raise RuntimeError("oracle go_dual_contour failed: %d" % n)
    tris = np.empty((max(n, 1), 3, 3), dtype=np.float32)
    lib().go_dual_contour(C.byref(tree.c), a, b, float(res), placer, tris.ctypes.data, n, st)
    return tris[:n], dict(levels=st[0], cubes=st[1], with_neighbors=st[2], evals=st[3])


def lsq_mgs64(A, b):
    """leastSquaresMGS64 (dual_contour_vertexplacement.go:148-223): float64 modified Gram-Schmidt solve of the K x 3
    system A x = b given in float32; returns x as float32[3]."""
    A = np.ascontiguousarray(A, dtype=np.float32).reshape(-1, 3)
    b = np.ascontiguousarray(b, dtype=np.float32).reshape(-1)
    assert len(A) == len(b)
    x = np.zeros(3, dtype=np.float32)
    f32p = C.POINTER(C.c_float)
    rc = lib().go_lsq_mgs64(len(A), A.ctypes.data_as(f32p), b.ctypes.data_as(f32p), x.ctypes.data_as(f32p))
    if rc != 0:
        raise RuntimeError("oracle go_lsq_mgs64: too many rows")
    return x


def dual_contour_quad_keys(tree, bbmin, bbmax, res, placer=DC_LSQ):
    """BFS cube key of the cube that emitted each quad of dual_contour()'s mesh (uint32, one per two triangles)."""
    a = (C.c_float * 3)(*[float(v) for v in bbmin])
    b = (C.c_float * 3)(*[float(v) for v in bbmax])
    n = lib().go_dual_contour(C.byref(tree.c), a, b, float(res), placer, None, 0, None)
    tris = np.empty((max(n, 1), 3, 3), dtype=np.float32)
    keys = np.zeros(max(n // 2, 1), dtype=np.uint32)
    lib().go_dual_contour_ex(C.byref(tree.c), a, b, float(res), placer, tris.ctypes.data, n, None, keys.ctypes.data)
    return keys[:n // 2]


def flat_lattice(bbmin, bbmax, res):
    lat = GoLattice()
    a = (C.c_float * 3)(*[float(v) for v in bbmin])
    b = (C.c_float * 3)(*[float(v) for v in bbmax])
    if lib().go_flat_lattice(a, b, float(res), C.byref(lat)) != 0:
        raise RuntimeError("resolution not fine enough for marching cubes")
    return lat


def octree_levels(bbmin, bbmax, res):
    a = (C.c_float * 3)(*[float(v) for v in bbmin])
    b = (C.c_float * 3)(*[float(v) for v in bbmax])
    return lib().go_octree_levels(a, b, float(res))


def flat_eval_grid(tree, lat, nthreads=1, batch=4096):
    """FlatRenderer.evalGrid: returns (grid float32[nz+1, ny+1, nx+1], evaluations)."""
    nx, ny, nz = lat.n
    grid = np.empty((nz + 1, ny + 1, nx + 1), dtype=np.float32)
    ev = lib().go_flat_eval_grid(C.byref(tree.c), C.byref(lat), grid.ctypes.data, nthreads, batch)
    if ev < 0:
        raise RuntimeError("oracle go_flat_eval_grid failed: %d" % ev)
    return grid, int(ev)


def flat_eval_planes(tree, lat, k0, k1, nthreads=1, batch=4096):
    """Corner planes [k0, k1) of the lattice: float32[k1-k0, ny+1, nx+1], positions as in the whole-lattice sweep."""
    nx, ny, nz = lat.n
    out = np.empty((k1 - k0, ny + 1, nx + 1), dtype=np.float32)
    ev = lib().go_flat_eval_planes(C.byref(tree.c), C.byref(lat), int(k0), int(k1), out.ctypes.data, nthreads, batch)
    if ev < 0:
        raise RuntimeError("oracle go_flat_eval_planes failed: %d" % ev)
    return out


def flat_march_planes(lat, planes, cz0, cz1, blockmask=None, want_cases=False):
    """FlatRenderer.ReadTriangles restricted to cell layers [cz0, cz1), given only the corner planes [cz0, cz1] of the
    lattice (flat_eval_planes(tree, lat, cz0, cz1 + 1)). Returns (triangles, cases[cz1-cz0, ny, nx] or None)."""
    nx, ny, nz = lat.n
    planes = np.ascontiguousarray(planes, dtype=np.float32)
    assert planes.shape == (cz1 - cz0 + 1, ny + 1, nx + 1)
    sz = (nx + 1) * (ny + 1) * 4
    gp = planes.ctypes.data - cz0 * sz  # go_flat_march_slab indexes planes absolutely and touches [cz0, cz1] only
    cases = np.empty((cz1 - cz0, ny, nx), dtype=np.uint8) if want_cases else None
    cp = (cases.ctypes.data - cz0 * nx * ny) if cases is not None else None
    mp = blockmask.ctypes.data if blockmask is not None else None
    mw = _mask_width(lat, blockmask)
    n = lib().go_flat_march_slab_w(C.byref(lat), gp, None, 0, cp, mp, mw, cz0, cz1)
    tris = np.empty((max(n, 1), 3, 3), dtype=np.float32)
    n = lib().go_flat_march_slab_w(C.byref(lat), gp, tris.ctypes.data, n, cp, mp, mw, cz0, cz1)
    return tris[:n], cases


def _mask_width(lat, blockmask):
    """Cube width (cells) a prune mask describes, from its shape: 4 (plan ending with level 3) or 2 (ending with level 2)."""
    if blockmask is None:
        return 4
    nx, ny, nz = lat.n
    for w in (4, 2):
        if tuple(blockmask.shape) == ((nz + w - 1) // w, (ny + w - 1) // w, (nx + w - 1) // w):
            assert blockmask.dtype == np.uint8 and blockmask.flags["C_CONTIGUOUS"]
            return w
    raise ValueError("prune mask shape %r fits neither 4-cell nor 2-cell cubes of the lattice" % (blockmask.shape,))


def octree_prune_mask(tree, lat):
    nx, ny, nz = lat.n
    mask = np.empty(((nz + 3) // 4, (ny + 3) // 4, (nx + 3) // 4), dtype=np.uint8)
    kept = lib().go_octree_prune_mask(C.byref(tree.c), C.byref(lat), mask.ctypes.data)
    if kept < 0:
        raise RuntimeError("oracle go_octree_prune_mask failed: %d" % kept)
    return mask, int(kept)


def octree_prune_plan(tree, lat, levels):
    """Coarse-to-fine prune (gsdf_prune_plan): levels = [(level, margin), ...] ending with level 3, or with level 3 followed
    by level 2. Returns the mask of the last level (4-cell or 2-cell cubes), its kept count and the number of cube centres
    evaluated."""
    nx, ny, nz = lat.n
    w = 1 << (int(levels[-1][0]) - 1)
    mask = np.empty(((nz + w - 1) // w, (ny + w - 1) // w, (nx + w - 1) // w), dtype=np.uint8)
    lv = (C.c_int * len(levels))(*[int(l) for l, _ in levels])
    mg = (C.c_float * len(levels))(*[float(m) for _, m in levels])
    ev = C.c_int64()
    kept = lib().go_octree_prune_plan(C.byref(tree.c), C.byref(lat), len(levels), lv, mg, mask.ctypes.data, C.byref(ev))
    if kept < 0:
        raise RuntimeError("oracle go_octree_prune_plan failed: %d" % kept)
    return mask, int(kept), int(ev.value)


def flat_march(lat, grid, want_cases=False, blockmask=None, max_tris=None, cz_range=None):
    """FlatRenderer.ReadTriangles sweep: returns (triangles (n,3,3), cases or None). cz_range restricts the sweep to
    one Z-slab of cell layers (grid / cases / blockmask still describe the whole lattice)."""
    nx, ny, nz = lat.n
    cz0, cz1 = cz_range if cz_range is not None else (0, nz)
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    cases = np.empty((nz, ny, nx), dtype=np.uint8) if want_cases else None
    mp = blockmask.ctypes.data if blockmask is not None else None
    cp = cases.ctypes.data if cases is not None else None
    mw = _mask_width(lat, blockmask)
    if max_tris is None:
        max_tris = lib().go_flat_march_slab_w(C.byref(lat), grid.ctypes.data, None, 0, cp, mp, mw, cz0, cz1)
    tris = np.empty((max(max_tris, 1), 3, 3), dtype=np.float32)
    n = lib().go_flat_march_slab_w(C.byref(lat), grid.ctypes.data, tris.ctypes.data, max_tris, cp, mp, mw, cz0, cz1)
    return tris[:min(n, max_tris)], cases


def stl_write(tris):
    tris = np.ascontiguousarray(tris, dtype=np.float32)
    n = tris.size // 9
    buf = np.empty(84 + 50 * max(n, 0), dtype=np.uint8)
    nb = lib().go_stl_write(tris.ctypes.data, n, buf.ctypes.data)
    if nb < 0:
        raise ValueError("empty triangle slice")
    return buf[:nb].tobytes()


def stl_read(data):
    arr = np.frombuffer(data, dtype=np.uint8)
    n = lib().go_stl_read(arr.ctypes.data, arr.size, None, 0)
    if n < 0:
        raise ValueError("bad STL")
    tris = np.empty((n, 3, 3), dtype=np.float32)
    lib().go_stl_read(arr.ctypes.data, arr.size, tris.ctypes.data, n)
    return tris
