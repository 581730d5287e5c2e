"""gsdf_b200 -- B200 (sm_100a) backend for the SDF evaluate + mesh hot path of soypat/gsdf.

Sub-modules mirror the reference's packages on that path:
    gsdf_b200.gsdf      gsdf.Builder, forge/threads, example scenes (host-side tree construction)
    gsdf_b200.gleval    gleval.SDF3 / SDF2 evaluators on the GPU
    gsdf_b200.glrender  Renderer / RenderAll / WriteBinarySTL / image evaluation
    gsdf_b200.gsdfaux   RenderShader3D driver (the caller of the hot path)
    gsdf_b200.slab      Z-slab partition helpers for one process per GPU

Importing the package loads gsdf_b200/libgsdfb200.so and fails loudly if it has not been built.
"""
from . import _lib  # noqa: F401  (raises ImportError with build instructions when the .so is missing)
from . import gsdf, gleval, glrender, gsdfaux, slab  # noqa: F401
from ._lib import GsdfError, lib  # noqa: F401


def version():
    return lib.gsdf_version().decode()


def device_count():
    return lib.gsdf_device_count()


def set_device(i):
    _lib.check(lib.gsdf_set_device(int(i)))
