"""Mirror of the reference's `glrender` package on the hot path: Renderer.ReadTriangles (glrender/glrender.go:11-13),
RenderAll (:17-36), NewOctreeRenderer (octreerenderer.go:45), FlatRenderer (flatrenderer.go:17), WriteBinarySTL /
ReadBinarySTL (stl.go:15,175) and ImageRendererSDF2's evaluation (image.go:76-105), all on the B200.

Triangles are numpy float32 arrays of shape (n, 3, 3): ms3.Triangle = [3]ms3.Vec.
"""
import ctypes as C
import io as _io
import struct

import numpy as np

from . import _lib
from ._lib import lib, check, GsdfError, Lattice, PrunePlan

marchingCubesMaxTriangles = 5  # marchcubes.go:11


class EOF(Exception):
    """io.EOF"""


class ErrShortBuffer(GsdfError):
    """io.ErrShortBuffer"""


def lattice_from_bounds(bbmin, bbmax, res):
    """FlatRenderer.Reset's lattice (flatrenderer.go:47-56)."""
    lat = Lattice()
    mn = (C.c_float * 3)(*[float(v) for v in bbmin])
    mx = (C.c_float * 3)(*[float(v) for v in bbmax])
    check(lib.gsdf_lattice_from_bounds(mn, mx, float(res), C.byref(lat)))
    return lat


class _Renderer:
    """Common part of the two renderers: the whole slab is meshed on the device when the renderer is created
    (or Reset); ReadTriangles streams the result out in FlatRenderer order."""
    _flags = 0

    def __init__(self, sdf, cubeResolution, evalBufferSize=0, numParallel=1, cz_range=None, keep_cases=False, keep_grid=False, stage_timing=False,
                 prune=None):
        self._h = None
        self.Reset(sdf, cubeResolution, cz_range=cz_range, keep_cases=keep_cases, keep_grid=keep_grid, stage_timing=stage_timing, prune=prune)

    def Reset(self, sdf, cubeResolution, evalBufferSize=0, numParallel=1, cz_range=None, keep_cases=False, keep_grid=False, stage_timing=False,
              prune=None):
        """prune (Octree only): None = the default plan (level 3 with margin 1.25, coarse levels in front on large
        lattices); "literal" = the reference's rule at every level-3 cube (margin 1); or an explicit plan
        [(level, margin), ...] ending with level 3 or with levels 3, 2 (include/gsdf_b200.h, gsdf_prune_plan)."""
        if not (cubeResolution > 0):
            raise GsdfError(_lib.EINVAL, "invalid renderer cube resolution")  # flatrenderer.go:38, octreerenderer.go:73
        self.Close()
        self.sdf = sdf
        mn, mx = sdf.Bounds()
        self.lat = lattice_from_bounds(mn, mx, cubeResolution)
        cz0, cz1 = cz_range if cz_range is not None else (0, self.lat.n[2])
        self.cz0, self.cz1 = int(cz0), int(cz1)
        flags = self._flags | (_lib.MESH_KEEP_CASES if keep_cases else 0) | (_lib.MESH_KEEP_GRID if keep_grid else 0) | \
            (_lib.MESH_STAGE_TIMING if stage_timing else 0)
        self.flags = flags
        self.plan = None
        h = C.c_void_p()
        if prune is not None and not (self._flags & _lib.MESH_PRUNE):
            raise GsdfError(_lib.EINVAL, "prune plans belong to the Octree renderer")
        if prune == "literal":
            flags |= _lib.MESH_PRUNE_LITERAL
            self.flags = flags
            prune = None
        if prune is not None:
            self.plan = PrunePlan.make([l for l, _ in prune], [m for _, m in prune])
            check(lib.gsdf_mesh_begin_plan(sdf._h, C.byref(self.lat), self.cz0, self.cz1, flags, C.byref(self.plan), C.byref(h)))
        else:
            if flags & _lib.MESH_PRUNE:
                self.plan = PrunePlan()
                check(lib.gsdf_prune_plan_default(C.byref(self.lat), flags, C.byref(self.plan)))
            check(lib.gsdf_mesh_begin(sdf._h, C.byref(self.lat), self.cz0, self.cz1, flags, C.byref(h)))
        self._h = h

    def Plan(self):
        """The prune plan in use: [(level, margin), ...], coarse to fine ([] for the FlatRenderer)."""
        return self.plan.levels() if self.plan is not None else []

    def Rerun(self):
        """Mesh the same slab again into the same device buffers (timing loops)."""
        check(lib.gsdf_mesh_rerun(self._h))

    def Rebind(self, sdf):
        """Keep lattice and device buffers, evaluate another SDF3CUDA on the next Rerun."""
        check(lib.gsdf_mesh_set_program(self._h, sdf._h))
        self.sdf = sdf

    def Close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.gsdf_mesh_destroy(h)

    __del__ = Close

    def ReadTriangles(self, dst, userData=None):
        """Renderer.ReadTriangles(dst []ms3.Triangle) -> n. Raises EOF when nothing is left (Go returns n, io.EOF)."""
        if dst.dtype != np.float32 or not dst.flags.c_contiguous:
            raise GsdfError(_lib.EINVAL, "dst must be C-contiguous float32 (n,3,3)")
        cap = dst.size // 9
        if cap < marchingCubesMaxTriangles:
            raise ErrShortBuffer(_lib.ESHORT, "short buffer")  # flatrenderer.go:187
        n = check(lib.gsdf_mesh_read(self._h, C.c_void_p(dst.ctypes.data), cap))
        if n == 0:
            raise EOF()
        return int(n)

    def _stats(self):
        e, p, t = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib.gsdf_mesh_stats(self._h, C.byref(e), C.byref(p), C.byref(t)))
        return e.value, p.value, t.value

    def Evaluations(self):
        return self._stats()[0]

    def TotalPruned(self):
        """Octree.TotalPruned() (octreerenderer.go:66): unit cubes inside pruned level-3 cubes."""
        return self._stats()[1]

    def NumTriangles(self):
        return self._stats()[2]

    def Timings(self):
        """Device milliseconds of the last run. Per-stage entries are only filled by a renderer created with
        stage_timing=True (eager launches with events between stages); otherwise reruns replay one CUDA graph and only
        total_ms is measured."""
        ms = (C.c_float * 5)()
        check(lib.gsdf_mesh_timings(self._h, ms))
        return dict(prune_ms=ms[0], eval_ms=ms[1], classify_ms=ms[2], emit_ms=ms[3], total_ms=ms[4])

    def Cases(self):
        nx, ny = self.lat.n[0], self.lat.n[1]
        out = np.empty((self.cz1 - self.cz0, ny, nx), dtype=np.uint8)
        check(lib.gsdf_mesh_cases(self._h, C.c_void_p(out.ctypes.data), out.size))
        return out

    def Grid(self):
        nx, ny = self.lat.n[0], self.lat.n[1]
        out = np.empty((self.cz1 - self.cz0 + 1, ny + 1, nx + 1), dtype=np.float32)
        check(lib.gsdf_mesh_grid(self._h, C.c_void_p(out.ctypes.data), out.size))
        return out

    def AllTriangles(self):
        """All triangles of the slab in one device->host copy."""
        nt = self.NumTriangles()
        out = np.empty((nt, 3, 3), dtype=np.float32)
        got = 0
        while got < nt:
            n = check(lib.gsdf_mesh_read(self._h, C.c_void_p(out[got:].ctypes.data), max(nt - got, 5)))
            if n == 0:
                break
            got += n
        return out[:got]

    def STLBytes(self):
        """WriteBinarySTL of the slab's triangles, packed on the device (one host write)."""
        nt = self.NumTriangles()
        buf = np.empty(84 + 50 * nt, dtype=np.uint8)
        n = check(lib.gsdf_mesh_stl(self._h, C.c_void_p(buf.ctypes.data), buf.size))
        return buf[:n].tobytes()


class SlabPipeline:
    """One lattice split into `nslabs` Z-slabs on ONE device, each with its own renderer: the device->host copy of
    slab i overlaps the kernels of slab i+1 (gsdf_mesh_read_async). Output = the slabs' triangles in slab order =
    FlatRenderer cell order, identical to a single renderer's. The same split across ranks is the multi-GPU layout."""

    def __init__(self, sdf, cubeResolution, nslabs=3, prune=True):
        from . import slab as _slab
        mn, mx = sdf.Bounds()
        lat = lattice_from_bounds(mn, mx, cubeResolution)
        cuts = _slab.slab_cuts(lat.n[2], nslabs)
        cls = Octree if prune else FlatRenderer
        self.parts = [cls(sdf, cubeResolution, cz_range=(a, b)) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        self.lat = lat

    def NumTriangles(self):
        return sum(p.NumTriangles() for p in self.parts)

    def RenderToHost(self, dst):
        """Re-render every slab and stream its triangles into dst (float32 (n,3,3), ideally pinned). Returns n.

        Steady state is fully asynchronous: every slab's render is enqueued (gsdf_mesh_rerun_begin) followed by the copy
        of as many triangles as that slab produced LAST time (gsdf_mesh_read_prefix_async), so slab i's copy runs under
        slab i+1's kernels and the host never waits for a count in between. The counts are checked afterwards; if any
        slab produced a different number (another tree was uploaded), the slabs are re-read at the right offsets."""
        flat = dst.reshape(-1)
        cap = flat.size // 9
        spec = getattr(self, "_last_counts", None)
        if spec is not None and sum(spec) <= cap:
            off = 0
            for p, n in zip(self.parts, spec):
                check(lib.gsdf_mesh_rerun_begin(p._h))
                if n:
                    got = check(lib.gsdf_mesh_read_prefix_async(p._h, C.c_void_p(flat[9 * off:].ctypes.data), n))
                    if got != n:
                        spec = None  # the slab's buffer shrank below the speculated count: fall through to the re-read
                        break
                off += n
            for p in self.parts:
                check(lib.gsdf_mesh_rerun_end(p._h))
            for p in self.parts:
                check(lib.gsdf_mesh_wait(p._h))
            counts = [p.NumTriangles() for p in self.parts]
            if spec is not None and counts == list(spec):
                return sum(counts)
        else:
            for p in self.parts:
                check(lib.gsdf_mesh_rerun(p._h))
            counts = [p.NumTriangles() for p in self.parts]
        # first call, or the speculation missed: read every slab at its true offset
        self._last_counts = counts
        if sum(counts) > cap:
            raise GsdfError(_lib.ESHORT, "destination holds %d triangles, the render produced %d" % (cap, sum(counts)))
        got = 0
        for p, n in zip(self.parts, counts):
            if n:
                k = check(lib.gsdf_mesh_read_async(p._h, C.c_void_p(flat[9 * got:].ctypes.data), max(n, 5)))
                assert k == n
            got += n
        for p in self.parts:
            check(lib.gsdf_mesh_wait(p._h))
        return got

    def Close(self):
        for p in self.parts:
            p.Close()


class Octree(_Renderer):
    """glrender.Octree: marching cubes with octree cube pruning (octreerenderer.go:15-284), coarse to fine down to the
    level-3 cubes (4 cells). See Reset for the prune plans."""
    _flags = _lib.MESH_PRUNE


class MultiRenderer:
    """One lattice meshed by several GPUs (or several pipelined Z-slabs of one GPU) from this process: gsdf_multi_*
    (include/gsdf_b200.h), the device analogue of FlatRenderer.evalGrid's goroutine split (flatrenderer.go:103-141).
    Satisfies the Renderer contract (ReadTriangles / RenderAll) on the last render; RenderToHost re-renders and delivers
    every triangle in FlatRenderer order into one host buffer."""

    def __init__(self, shader, cubeResolution, devices=(0,), slabs_per_device=1, prune=True, literal=False):
        if not (cubeResolution > 0):
            raise GsdfError(_lib.EINVAL, "invalid renderer cube resolution")
        self._h = None
        self.shader = shader
        mn, mx = shader.Bounds()
        self.lat = lattice_from_bounds(mn, mx, cubeResolution)
        f = shader.bld.flatten(shader)
        aux = np.ascontiguousarray(f["aux"], dtype=np.float32)
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        flags = (_lib.MESH_PRUNE if prune else 0) | (_lib.MESH_PRUNE_LITERAL if literal else 0)
        h = C.c_void_p()
        check(lib.gsdf_multi_begin(len(devices), devs, int(slabs_per_device), f["blob"], len(f["blob"]),
                                   aux.ctypes.data_as(C.POINTER(C.c_float)), aux.size, C.byref(self.lat), flags, C.byref(h)))
        self._h = h

    def Update(self, shader):
        """Re-flatten `shader` and hand it to every device (uploaded at the start of the next render)."""
        f = shader.bld.flatten(shader)
        aux = np.ascontiguousarray(f["aux"], dtype=np.float32)
        check(lib.gsdf_multi_update(self._h, f["blob"], len(f["blob"]), aux.ctypes.data_as(C.POINTER(C.c_float)), aux.size))
        self.shader = shader

    def Specialize(self):
        """gsdf_multi_specialize: run-time compiled kernels for this tree on every device (see gleval SDF3CUDA.Specialize).
        True when in use, False when run-time compilation is not available."""
        rc = lib.gsdf_multi_specialize(self._h)
        if rc == _lib.EUNSUPPORTED:
            return False
        check(rc)
        return True

    def UpdateBlob(self, blob, aux):
        check(lib.gsdf_multi_update(self._h, blob, len(blob), aux.ctypes.data_as(C.POINTER(C.c_float)), aux.size))

    def Rebalance(self, rounds=2):
        """Re-cut the Z-slabs by the evaluations each executed in its last render (gsdf_multi_rebalance)."""
        return int(check(lib.gsdf_multi_rebalance(self._h, int(rounds))))

    def Render(self):
        """Render without read-back (device timing)."""
        return int(check(lib.gsdf_multi_render(self._h, None, 0)))

    def RenderToHost(self, dst):
        """Render and deliver all triangles into dst (float32, capacity >= the count; pinned memory is written by DMA
        directly). Returns the triangle count."""
        if dst.dtype != np.float32 or not dst.flags.c_contiguous:
            raise GsdfError(_lib.EINVAL, "dst must be C-contiguous float32")
        return int(check(lib.gsdf_multi_render(self._h, C.c_void_p(dst.ctypes.data), dst.size // 9)))

    def ReadTriangles(self, dst, userData=None):
        if dst.dtype != np.float32 or not dst.flags.c_contiguous:
            raise GsdfError(_lib.EINVAL, "dst must be C-contiguous float32 (n,3,3)")
        cap = dst.size // 9
        if cap < marchingCubesMaxTriangles:
            raise ErrShortBuffer(_lib.ESHORT, "short buffer")
        n = check(lib.gsdf_multi_read(self._h, C.c_void_p(dst.ctypes.data), cap))
        if n == 0:
            raise EOF()
        return int(n)

    def _stats(self):
        e, p, t, ms = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_float()
        check(lib.gsdf_multi_stats(self._h, C.byref(e), C.byref(p), C.byref(t), C.byref(ms)))
        return e.value, p.value, t.value, ms.value

    def Evaluations(self):
        return self._stats()[0]

    def TotalPruned(self):
        return self._stats()[1]

    def NumTriangles(self):
        return self._stats()[2]

    def DeviceMs(self):
        return self._stats()[3]

    def Timeline(self):
        """Host-clock microseconds of the last RenderToHost (gsdf_multi_timeline): dict(enqueued, slabs=[(count seen, copy
        enqueued)], delivered)."""
        buf = (C.c_double * 8200)()
        n = check(lib.gsdf_multi_timeline(self._h, buf, 8200))
        v = [float(x) for x in buf[:n]]
        return dict(enqueued=v[0], slabs=[(v[1 + 2 * j], v[2 + 2 * j]) for j in range((n - 2) // 2)], delivered=v[-1])

    def Slabs(self):
        """(cuts, devices): nslabs+1 cell-layer cuts and the device of each slab."""
        cuts = (C.c_int32 * 4097)()
        devs = (C.c_int32 * 4096)()
        n = check(lib.gsdf_multi_slabs(self._h, cuts, devs, 4096))
        return [int(c) for c in cuts[:n + 1]], [int(d) for d in devs[:n]]

    def AllTriangles(self):
        nt = self.NumTriangles()
        out = np.empty((nt, 3, 3), dtype=np.float32)
        got = 0
        check(lib.gsdf_multi_rewind(self._h))
        while got < nt:
            n = check(lib.gsdf_multi_read(self._h, C.c_void_p(out[got:].ctypes.data), max(nt - got, 5)))
            if n == 0:
                break
            got += n
        return out[:got]

    def STLBytes(self):
        nt = self.NumTriangles()
        buf = np.empty(84 + 50 * nt, dtype=np.uint8)
        n = check(lib.gsdf_multi_stl(self._h, C.c_void_p(buf.ctypes.data), buf.size))
        return buf[:n].tobytes()

    def Close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.gsdf_multi_destroy(h)

    __del__ = Close


def pinned_empty(shape, dtype=np.float32):
    """A numpy array in page-locked host memory from gsdf_host_alloc: transfers into it run by DMA without staging."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.gsdf_host_alloc(max(n, 1))
    if not p:
        raise GsdfError(_lib.ENOMEM, _lib.last_error())
    buf = (C.c_uint8 * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    import weakref
    weakref.finalize(buf, lib.gsdf_host_free, p)  # buf is arr's base: freed when the last view goes
    return arr


class FlatRenderer(_Renderer):
    """glrender.FlatRenderer: every lattice corner evaluated once (flatrenderer.go:17-256)."""
    _flags = 0


def octree_levels(bbmin, bbmax, res):
    """makeICube's level count (octreerenderer.go:222-235); raises for a resolution too coarse to march."""
    mn = (C.c_float * 3)(*[float(v) for v in bbmin])
    mx = (C.c_float * 3)(*[float(v) for v in bbmax])
    return check(lib.gsdf_octree_levels(mn, mx, float(res)))


def NewOctreeRenderer(s, cubeResolution, evalBufferSize=64, **kw):
    if evalBufferSize < 64:
        raise GsdfError(_lib.EINVAL, "bad octree eval buffer size")  # octreerenderer.go:46
    octree_levels(*s.Bounds(), cubeResolution)  # early error check, octreerenderer.go:50-53
    return Octree(s, cubeResolution, evalBufferSize, **kw)


def NewFlatRenderer(s, cubeResolution, evalBufferSize=4096, numParallel=1, **kw):
    if evalBufferSize < 8:
        raise GsdfError(_lib.EINVAL, "flat renderer eval buffer size must be at least 8")  # flatrenderer.go:41
    if numParallel < 1:
        raise GsdfError(_lib.EINVAL, "flat renderer numParallel must be at least 1")       # flatrenderer.go:44
    return FlatRenderer(s, cubeResolution, evalBufferSize, numParallel, **kw)


def RenderAll(r, userData=None):
    """glrender.RenderAll (glrender.go:17-36): 4096-triangle buffer until io.EOF."""
    startSize = 4096
    buf = np.empty((startSize, 3, 3), dtype=np.float32)
    parts = []
    while True:
        try:
            n = r.ReadTriangles(buf, userData)
        except EOF:
            break
        parts.append(buf[:n].copy())
    return np.concatenate(parts) if parts else np.zeros((0, 3, 3), np.float32)


def WriteBinarySTL(w, model):
    """glrender.WriteBinarySTL(w io.Writer, model []ms3.Triangle) (stl.go:15-62). Returns bytes written."""
    model = np.ascontiguousarray(model, dtype=np.float32)
    n = model.size // 9
    if n == 0:
        raise GsdfError(_lib.EEMPTY, "empty triangle slice")  # stl.go:16
    buf = np.empty(84 + 50 * n, dtype=np.uint8)
    nb = check(lib.gsdf_stl_pack(C.c_void_p(model.ctypes.data), n, C.c_void_p(buf.ctypes.data), buf.size))
    w.write(buf[:nb].tobytes())
    return int(nb)


def ReadBinarySTL(r):
    """glrender.ReadBinarySTL (stl.go:175-225): returns the triangles (n,3,3); normals are not returned."""
    data = r.read()
    if len(data) < 84:
        raise GsdfError(_lib.EINVAL, "encountered EOF while reading STL header")
    (count,) = struct.unpack_from("<I", data, 80)
    if count == 0:
        raise GsdfError(_lib.EINVAL, "STL header indicates 0 triangles present")
    if len(data) < 84 + 50 * count:
        raise GsdfError(_lib.EINVAL, "%d/%d STL triangles read: unexpected EOF" % ((len(data) - 84) // 50, count))
    rec = np.frombuffer(data, dtype=np.uint8, count=50 * count, offset=84).reshape(count, 50)
    tris = rec[:, 12:48].copy().view(np.float32).reshape(count, 3, 3)
    normals = rec[:, 0:12].copy().view(np.float32).reshape(count, 3)
    # stlTriangle.validate (stl.go:129-150): inf/NaN and degenerate triangles are errors; a stored normal that is
    # neither +n nor -n of the vertices (tolerance 5e-2) is counted, more than 10,000 of them is an error (:207-214).
    if not np.isfinite(normals).all():
        raise GsdfError(_lib.EINVAL, "inf/NaN STL triangle normal")
    if not np.isfinite(tris).all():
        raise GsdfError(_lib.EINVAL, "inf/NaN STL triangle vertex")
    v = tris.astype(np.float64) * 10.0
    calc = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
    ln = np.linalg.norm(calc, axis=1)
    degenerate = ln <= 1e-12
    if degenerate.any():
        raise GsdfError(_lib.EINVAL, "%d/%d STL triangles read: triangle is degenerate" % (int(np.argmax(degenerate)) + 1, count))
    calc /= ln[:, None]
    bad = ~((np.abs(calc - normals) <= 5e-2).all(axis=1) | (np.abs(-calc - normals) <= 5e-2).all(axis=1))
    if int(bad.sum()) > 10000:
        raise GsdfError(_lib.EINVAL, "got too many normal vector mismatches (%d)" % int(bad.sum()))
    return tris


def ImageEvaluateSDF2(sdf2, width, height):
    """The evaluation ImageRendererSDF2.Render performs (image.go:76-105): distances at pixel centres, row 0 at
    Bounds().Max.Y. Returns float32 (height, width)."""
    mn, mx = sdf2.Bounds()
    out = np.empty((height, width), dtype=np.float32)
    a = (C.c_float * 2)(float(mn[0]), float(mn[1]))
    b = (C.c_float * 2)(float(mx[0]), float(mx[1]))
    check(lib.gsdf_image_eval2(sdf2._h, a, b, int(width), int(height), C.c_void_p(out.ctypes.data)))
    return out


class ImageRendererSDF2:
    """glrender.ImageRendererSDF2 (image.go:20-118) with the colour conversion applied on the device.

    `conversion` is None (black inside / white outside / red NaN, image.go:50-61) or a _lib.ColorConv made by
    gsdfaux.ColorConversionLinearGradient / ColorConversionInigoQuilez -- the reference takes a Go closure; the
    closures gsdfaux provides are data here so the kernel that evaluates a pixel also colours it."""

    def __init__(self, evalBufferSize, conversion=None):
        if evalBufferSize < 4096:
            raise GsdfError(_lib.EINVAL, "too small evaluation buffer size")  # image.go:46-48
        if conversion is not None and not isinstance(conversion, _lib.ColorConv):
            raise GsdfError(_lib.EINVAL, "conversion must be None or a gsdfaux.ColorConversion* value")
        self.evalBufferSize, self.conv = int(evalBufferSize), conversion

    def Render(self, sdf2, img, userData=None):
        """Fills img, a uint8 array (height, width, 4) in image.RGBA.Pix order. userData is ignored (gpu.go:82)."""
        if not (isinstance(img, np.ndarray) and img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 4 and img.flags.c_contiguous):
            raise GsdfError(_lib.EINVAL, "img must be a C-contiguous uint8 array (height, width, 4)")
        h, w = img.shape[:2]
        if self.evalBufferSize < w:  # image.go:80-82
            raise GsdfError(_lib.EINVAL, "require evaluation buffer (%d) to be at least of length of image rows (%d)" % (self.evalBufferSize, w))
        mn, mx = sdf2.Bounds()
        a = (C.c_float * 2)(float(mn[0]), float(mn[1]))
        b = (C.c_float * 2)(float(mx[0]), float(mx[1]))
        check(lib.gsdf_image_render2(sdf2._h, a, b, int(w), int(h), C.byref(self.conv) if self.conv is not None else None,
                                     C.c_void_p(img.ctypes.data)))


def NewImageRendererSDF2(evalBufferSize, conversion=None):
    return ImageRendererSDF2(evalBufferSize, conversion)


class DualContourLeastSquares:
    """glrender.DualContourLeastSquares (dual_contour_vertexplacement.go:16-23): QEF vertex placement."""

    def __init__(self, Chiseled=False):
        self.Chiseled = bool(Chiseled)

    @property
    def _kind(self):
        return _lib.DC_LEAST_SQUARES_CHISELED if self.Chiseled else _lib.DC_LEAST_SQUARES


class DualContourNaive:
    """Mean-of-crossings vertex placement (the DualContourer of dual_contour_test.go:355-389)."""
    _kind = _lib.DC_NAIVE


class DualContourRenderer:
    """glrender.DualContourRenderer (dual_contour.go:12-218): Reset(sdf, res, vertexPlacer) then RenderAll(dst)."""

    def __init__(self):
        self._h = None

    def Reset(self, sdf, res, vertexPlacer, userData=None, part=0, nparts=1):
        """part / nparts (1, 2, 4, 8): multi-GPU split by top-level octants; rank r passes part=r, nparts=world size and
        the ranks' RenderAll results concatenated in rank order equal the single-renderer mesh."""
        if vertexPlacer is None or not hasattr(vertexPlacer, "_kind"):
            raise GsdfError(_lib.EINVAL, "nil DualContourer argument to Reset")  # dual_contour.go:28-30
        self.Close()
        mn, mx = sdf.Bounds()
        a = (C.c_float * 3)(*[float(v) for v in mn])
        b = (C.c_float * 3)(*[float(v) for v in mx])
        h = C.c_void_p()
        check(lib.gsdf_dc_begin_part(sdf._h, a, b, float(res), int(vertexPlacer._kind), int(part), int(nparts), C.byref(h)))
        self._h, self._sdf = h, sdf

    def Rerun(self):
        check(lib.gsdf_dc_rerun(self._h))

    def Stats(self):
        st = (C.c_uint64 * 6)()
        check(lib.gsdf_dc_stats(self._h, st))
        return dict(levels=int(st[0]), cubes=int(st[1]), with_neighbors=int(st[2]), triangles=int(st[3]), evals=int(st[4]), device_ms=st[5] / 1000.0)

    def RenderAll(self, dst=None, userData=None):
        """Appends the mesh to dst (float32 (n,3,3) or None) and returns the result, like RenderAll(dst, userData)."""
        if self._h is None:
            raise GsdfError(_lib.EINVAL, "DualContourRenderer.RenderAll before Reset")
        n = self.Stats()["triangles"]
        out = np.empty((n, 3, 3), np.float32)
        if n:
            check(lib.gsdf_dc_read(self._h, C.c_void_p(out.ctypes.data), n))
        if dst is not None and len(dst):
            return np.concatenate([np.asarray(dst, np.float32).reshape(-1, 3, 3), out])
        return out

    def Close(self):
        h, self._h = self._h, None
        if h is not None and lib is not None:
            lib.gsdf_dc_destroy(h)

    def __del__(self):
        try:
            self.Close()
        except Exception:
            pass
