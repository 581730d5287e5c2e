"""The device interpreter's own source (gsdf_b200/csrc/interp.cuh + math32.cuh) compiled for the HOST with g++ and run
from the tests (tests/host_interp/: a stand-in <cuda_runtime.h> and a 40-line driver). Where tests/progsim.py is an
independent model of the interpreter, this is the interpreter: the same opcode bodies, stack handling and guard logic the
GPU executes, with the math32 helpers on their host branches (sqrtf, a / b -- the same IEEE operations as the device's
_rn intrinsics) and -ffp-contract=off in place of -fmad=false. Test infrastructure only."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host_interp")
OUT = os.path.join(SRC, "_build", "libhost_interp.so")
_lib = None
_libs = {}
EXT_OPS = {18, 19}  # GSDF_OP_ELLIPSE2D, GSDF_OP_BEZIERQ2D -> the EXT instantiation


def build(extra=(), out=None):
    global OUT
    saved = OUT
    if out:
        OUT = out
    try:
        return _build(extra)
    finally:
        OUT = saved


def _build(extra=()):
    deps = [os.path.join(SRC, "host_interp.cpp"), os.path.join(SRC, "shim", "cuda_runtime.h"),
            os.path.join(ROOT, "gsdf_b200", "csrc", "interp.cuh"), os.path.join(ROOT, "gsdf_b200", "csrc", "math32.cuh"),
            os.path.join(ROOT, "include", "gsdf_program.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps) or extra:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas", "-shared", "-fPIC",
                               "-I", os.path.join(SRC, "shim"), "-I", os.path.join(ROOT, "gsdf_b200", "csrc"), *extra, "-o", OUT, deps[0]])
    return OUT


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.host_interp_eval.restype = C.c_int
        L.host_interp_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        _lib = L
    return _lib


def lib_variant(define):
    """A build of the interpreter source with an extra -D (e.g. GSDF_RXY), kept beside the default one."""
    if define not in _libs:
        path = build(extra=("-D" + define,), out=os.path.join(SRC, "_build", "libhost_interp_%s.so" % define.lower()))
        L = C.CDLL(path)
        L.host_interp_eval.restype = C.c_int
        L.host_interp_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        _libs[define] = L
    return _libs[define]


def run(flat, pos, variant=None):
    """Distances of a flattened program (the dict gsdf.Builder.flatten returns) at pos ((n,3) or (n,2) float32)."""
    blob = flat["blob"]
    magic, ver, nchunks, dim, dstack, pstack, ninstr, _ = struct.unpack_from("<8I", blob, 0)
    assert magic == 0x46445347 and ver == 1
    chunks = np.frombuffer(blob, np.uint32, offset=32).copy()
    aux = np.ascontiguousarray(flat["aux"], np.float32)
    if aux.size == 0:
        aux = np.zeros(4, np.float32)
    pos = np.ascontiguousarray(pos, np.float32)
    assert pos.ndim == 2 and pos.shape[1] == dim
    ext, pc = 0, 0
    while pc < nchunks:
        op, ln = int(chunks[4 * pc]) & 0xff, (int(chunks[4 * pc]) >> 8) & 0xff
        ext |= op in EXT_OPS
        if op == 0:
            break
        pc += ln
    out = np.empty(len(pos), np.float32)
    rc = (lib_variant(variant) if variant else lib()).host_interp_eval(chunks.ctypes.data, aux.ctypes.data, dstack, pstack, dim, pos.ctypes.data, out.ctypes.data, len(pos), int(ext))
    assert rc == 0
    return out
