"""Mirror of the reference's `gleval` package boundary: gleval.SDF3 / gleval.SDF2 (gleval/gleval.go:15-37) evaluated
on the B200 through libgsdfb200.so.

    sdf = gleval.NewCUDASDF3(shader)       # like gleval.NewComputeGPUSDF3(source, bb, cfg)  gleval/gpu.go:35
    sdf.Evaluate(pos, dist)                # pos: (n,3) float32, dist: (n,) float32          gleval/gleval.go:21
    sdf.Bounds(); sdf.Evaluations()

Host numpy arrays are copied in and out inside the call (the reference-facing path); torch CUDA tensors are
evaluated in place in HBM with no copies.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check, GsdfError


class _ErrMismatchBufferLength(GsdfError):
    pass


errMismatchBufferLength = "position and distance buffer length mismatch"  # gleval/gleval.go:48
errEmptyBuffers = "empty buffers"                                          # gleval/gleval.go:47


def _is_torch_cuda(x):
    return hasattr(x, "is_cuda") and x.is_cuda


class _SDFCUDA:
    _dim = 3

    def __init__(self, shader):
        if shader.is2d != (self._dim == 2):
            raise GsdfError(_lib.EINVAL, "%s does not implement %dD evaluator" % ("Shader", self._dim))  # gleval/cpu.go:60-66
        self.shader = shader
        h = C.c_void_p()
        try:
            f = shader.bld.flatten(shader)  # host layer (libgsdfhost.so): tree -> node program
        except GsdfError as e:
            shader.bld.ClearErrors()  # reported through the exception: must not poison later shape construction on this builder
            raise e
        aux = np.ascontiguousarray(f["aux"], dtype=np.float32)
        rc = lib.gsdf_program_create(f["blob"], len(f["blob"]), aux.ctypes.data_as(C.POINTER(C.c_float)), aux.size, C.byref(h))
        if rc != 0:
            raise GsdfError(rc, _lib.last_error())
        self._h = h
        self._bounds = shader.Bounds()

    def Update(self, shader):
        """Re-flatten `shader` (same dimension) and upload it into this evaluator's device buffers."""
        f = shader.bld.flatten(shader)
        aux = np.ascontiguousarray(f["aux"], dtype=np.float32)
        check(lib.gsdf_program_update(self._h, f["blob"], len(f["blob"]), aux.ctypes.data_as(C.POINTER(C.c_float)), aux.size))
        self.shader = shader
        self._bounds = shader.Bounds()

    def Specialize(self):
        """Compile kernels specialised for this tree's instruction stream (gsdf_program_specialize: what constructing the GPU
        evaluator does in the reference, which compiles a GLSL shader per tree, gleval/gpu.go:35-54). Renderers built on this
        evaluator then use them for the lattice evaluation and the prune-centre passes: bit-identical results, faster.
        Returns True when the specialised kernels are in use, False when run-time compilation is not available here (the
        interpreter kernels keep running)."""
        rc = lib.gsdf_program_specialize(self._h)
        if rc == _lib.EUNSUPPORTED:
            return False
        check(rc)
        return True

    def Specialized(self):
        return bool(lib.gsdf_program_is_specialized(self._h))

    def Close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:  # lib can already be gone at interpreter shutdown
            lib.gsdf_program_destroy(h)

    __del__ = Close

    def Bounds(self):
        return self._bounds

    def Evaluations(self):
        """gleval Evaluations() counter (gleval/gpu.go:80)."""
        return int(lib.gsdf_program_evaluations(self._h))

    def Evaluate(self, pos, dist, userData=None):
        """gleval.SDF3.Evaluate / SDF2.Evaluate. userData is ignored, as the GL evaluator ignores it (gpu.go:82)."""
        d = self._dim
        if _is_torch_cuda(pos) or _is_torch_cuda(dist):
            import torch
            if not (_is_torch_cuda(pos) and _is_torch_cuda(dist)):
                raise GsdfError(_lib.EINVAL, "pos and dist must both be CUDA tensors")
            if pos.dtype != torch.float32 or dist.dtype != torch.float32 or not pos.is_contiguous() or not dist.is_contiguous():
                raise GsdfError(_lib.EINVAL, "pos/dist must be contiguous float32")
            n = pos.numel() // d
            if n != dist.numel():
                raise GsdfError(_lib.ELEN, errMismatchBufferLength)
            if n == 0:
                raise GsdfError(_lib.EEMPTY, errEmptyBuffers)
            fn = lib.gsdf_eval3_device if d == 3 else lib.gsdf_eval2_device
            # torch's default stream has handle 0, which the C ABI reads as "the handle's own stream": pass the legacy
            # default stream's explicit handle (cudaStreamLegacy = 1) so the kernel is ordered with the tensors' producers
            stream = torch.cuda.current_stream(pos.device).cuda_stream or 1
            check(fn(self._h, C.c_void_p(pos.data_ptr()), C.c_void_p(dist.data_ptr()), n, C.c_void_p(stream)))
            return
        if not (isinstance(pos, np.ndarray) and isinstance(dist, np.ndarray)):
            raise GsdfError(_lib.EINVAL, "pos/dist must be numpy arrays or torch CUDA tensors")
        if pos.dtype != np.float32 or dist.dtype != np.float32 or not pos.flags.c_contiguous or not dist.flags.c_contiguous:
            raise GsdfError(_lib.EINVAL, "pos/dist must be C-contiguous float32")
        n = pos.size // d
        if n != dist.size or pos.size != n * d:
            raise GsdfError(_lib.ELEN, errMismatchBufferLength)  # gleval/cpu.go:95-96
        if n == 0:
            raise GsdfError(_lib.EEMPTY, errEmptyBuffers)         # gleval/cpu.go:97-98
        fn = lib.gsdf_eval3 if d == 3 else lib.gsdf_eval2
        check(fn(self._h, C.c_void_p(pos.ctypes.data), C.c_void_p(dist.ctypes.data), n))


class SDF3CUDA(_SDFCUDA):
    _dim = 3


class SDF2CUDA(_SDFCUDA):
    _dim = 2


def NewCUDASDF3(shader):
    return SDF3CUDA(shader)


def NewCUDASDF2(shader):
    return SDF2CUDA(shader)


# ---------------------------------------------------------------------------------------------- special evaluators
class ComputeConfig:
    """gleval.ComputeConfig (gleval/gpu.go:64-73). InvocX (the GL work-group size) is validated like the reference does
    and otherwise unused: CTA sizes are fixed by the CUDA kernels."""

    def __init__(self, InvocX=1, ShaderObjects=None, CompileFlags=None):
        self.InvocX, self.ShaderObjects, self.CompileFlags = int(InvocX), ShaderObjects, CompileFlags


class _SpecialSDF2:
    """Common part of gleval.PolygonGPU / Lines2DGPU / DisplaceMulti2D (gleval/gpu.go:169-446): stand-alone gleval.SDF2
    evaluators the GL backend compiles one dedicated shader for. Here each is a one-node (or one-operator) program for
    the same interpreter kernel."""

    def __init__(self):
        self._sdf = None
        self._bld = None

    def _configure(self, cfg, build):
        if cfg is None or cfg.InvocX < 1:
            raise GsdfError(_lib.EINVAL, "zero or negative invocation size")  # errZeroInvoc, gleval/gpu.go
        from . import gsdf as _gsdf
        self._bld = _gsdf.Builder()
        self._sdf = SDF2CUDA(build(self._bld))

    def Evaluate(self, pos, dist, userData=None):
        if self._sdf is None:
            raise GsdfError(_lib.EINVAL, "evaluator used before Configure")
        self._sdf.Evaluate(pos, dist, userData)

    def Evaluations(self):
        return self._sdf.Evaluations() if self._sdf is not None else 0


class PolygonGPU(_SpecialSDF2):
    """gleval.PolygonGPU{Vertices} (gleval/gpu.go:169-258): direct polygon SDF (winding-number form of poly2D)."""

    def __init__(self, Vertices):
        super().__init__()
        self.Vertices = np.ascontiguousarray(Vertices, dtype=np.float32).reshape(-1, 2)

    def Configure(self, cfg):
        self._configure(cfg, lambda b: b.NewPolygon(self.Vertices))

    def Bounds(self):
        """gleval/gpu.go:181-191: componentwise min/max of the vertices."""
        return self.Vertices.min(axis=0), self.Vertices.max(axis=0)


class Lines2DGPU(_SpecialSDF2):
    """gleval.Lines2DGPU{Lines, Width} (gleval/gpu.go:260-352): union of line segments of a given width."""

    def __init__(self, Lines, Width):
        super().__init__()
        self.Lines = np.ascontiguousarray(Lines, dtype=np.float32).reshape(-1, 2, 2)
        self.Width = np.float32(Width)

    def Configure(self, cfg):
        self._configure(cfg, lambda b: b.NewLines2D(self.Lines, self.Width))

    def Bounds(self):
        """gleval/gpu.go:277-291: segment endpoints grown by Width/2."""
        off = self.Width / np.float32(2)
        pts = self.Lines.reshape(-1, 2)
        return (pts - off).min(axis=0).astype(np.float32), (pts + off).max(axis=0).astype(np.float32)


class DisplaceMulti2D(_SpecialSDF2):
    """gleval.DisplaceMulti2D{Displacements} (gleval/gpu.go:354-446): union of one 2-D element translated to many
    places. Configure(element_builder, cfg): element_builder(bld) must return the Shader2D built on the given Builder
    (the reference passes a Programmer + Shader2D, gpu.go:381)."""

    def __init__(self, Displacements):
        super().__init__()
        self.Displacements = np.ascontiguousarray(Displacements, dtype=np.float32).reshape(-1, 2)
        self._elem_bb = None

    def Configure(self, element_builder, cfg):
        def build(b):
            elem = element_builder(b)
            self._elem_bb = elem.Bounds()
            return b.TranslateMulti2D(elem, self.Displacements)
        self._configure(cfg, build)

    def Bounds(self):
        """gleval/gpu.go:372-379: union of the element box moved by every displacement."""
        mn, mx = self._elem_bb
        return (mn + self.Displacements.min(axis=0)).astype(np.float32), (mx + self.Displacements.max(axis=0)).astype(np.float32)


def NormalsCentralDiff(s, pos, normals, step, userData=None):
    """gleval.NormalsCentralDiff (gleval/gleval.go:53-108): un-normalised central differences, six Evaluate calls."""
    step = np.float32(step) * np.float32(0.5)
    if not (step > 0):
        raise GsdfError(_lib.EINVAL, "invalid step")
    if len(pos) != len(normals):
        raise GsdfError(_lib.EINVAL, "length of position must match length of normals")
    if s is None:
        raise GsdfError(_lib.EINVAL, "nil SDF3")
    if len(pos) == 0:
        raise GsdfError(_lib.EEMPTY, errEmptyBuffers)
    pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
    d1 = np.empty(len(pos), np.float32)
    d2 = np.empty(len(pos), np.float32)
    for dim in range(3):
        h = np.zeros(3, np.float32)
        h[dim] = step
        s.Evaluate(np.ascontiguousarray(pos + h), d1, userData)
        s.Evaluate(np.ascontiguousarray(pos - h), d2, userData)
        normals[:, dim] = d1 - d2
