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
        rc = lib.gsdfh_compile(shader.bld._h, shader.id, C.byref(h))
        if rc != 0:
            raise GsdfError(rc, shader.bld.Err() or _lib.last_error())
        self._h = h
        self._bounds = shader.Bounds()

    def Update(self, shader):
        """Re-flatten `shader` (same dimension) and upload it into this evaluator's device buffers."""
        f = shader.bld.flatten(shader)
        aux = np.ascontiguousarray(f["aux"], dtype=np.float32)
        check(lib.gsdf_program_update(self._h, f["blob"], len(f["blob"]), aux.ctypes.data_as(C.POINTER(C.c_float)), aux.size))
        self.shader = shader
        self._bounds = shader.Bounds()

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
            stream = torch.cuda.current_stream(pos.device).cuda_stream
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
