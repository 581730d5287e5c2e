"""Mirror of the reference's driver gsdfaux.RenderShader3D (gsdfaux/gsdfaux.go:63-241) for the CUDA backend: pick the
evaluator, pick the renderer, RenderAll, log what the reference logs, WriteBinarySTL. It is the CALLER of the hot
path, kept thin; bench.py measures the same sequence."""
import ctypes as C
import struct
import time
import zlib

import numpy as np

from . import _lib, gleval, glrender
from .gsdf import _hypot32
from ._lib import check, lib

Black, White = 0xff000000, 0xffffffff  # color.Black / color.White as R | G<<8 | B<<16 | A<<24


def RGBA(r, g, b, a=255):
    """color.RGBA{r,g,b,a} packed the way gsdf_colorconv takes colours."""
    return (int(r) & 255) | (int(g) & 255) << 8 | (int(b) & 255) << 16 | (int(a) & 255) << 24


def ColorConversionInigoQuilez(characteristicDistance):
    """gsdfaux.ColorConversionInigoQuilez (color.go:21-47) as data for the fused image kernel."""
    cc = _lib.ColorConv()
    check(lib.gsdf_colorconv_inigo_quilez(float(characteristicDistance), C.byref(cc)))
    return cc


def ColorConversionLinearGradient(gradientLength, c0, c1):
    """gsdfaux.ColorConversionLinearGradient (color.go:51-73); Black -> White selects the grayscale form (:52-54)."""
    cc = _lib.ColorConv()
    check(lib.gsdf_colorconv_linear_gradient(float(gradientLength), int(c0), int(c1), C.byref(cc)))
    return cc


def _png_bytes(img):
    """Minimal RGBA8 PNG encoder (image/png's role in RenderPNGFile, gsdfaux.go:289): filter 0, one IDAT."""
    h, w = img.shape[:2]
    raw = np.empty((h, 1 + 4 * w), np.uint8)
    raw[:, 0] = 0
    raw[:, 1:] = img.reshape(h, 4 * w)

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)) + chunk(b"IEND", b"")


def RenderPNGFile(filename, sdf, picHeight, colorConversion=None):
    """gsdfaux.RenderPNGFile (gsdfaux.go:267-296): width from the aspect ratio (float64 like the reference, :273-274),
    nil conversion -> Inigo Quilez with characteristic distance Diagonal/3 (:269-271). Returns the RGBA array."""
    mn, mx = sdf.Bounds()
    if colorConversion is None:
        sz = (mx - mn).astype(np.float32)
        colorConversion = ColorConversionInigoQuilez(_hypot32(sz[0], sz[1]) / np.float32(3))
    sz = (mx - mn).astype(np.float32)
    pixPerUnit = float(picHeight) / float(sz[1])
    picWidth = int(pixPerUnit * float(sz[0]))
    img = np.empty((int(picHeight), picWidth, 4), np.uint8)
    renderer = glrender.NewImageRendererSDF2(max(4096, picWidth), colorConversion)
    renderer.Render(sdf, img)
    with open(filename, "wb") as fp:
        fp.write(_png_bytes(img))
    return img


class RenderConfig:
    """gsdfaux.RenderConfig (gsdfaux.go:25-39). UseGPU selects the Octree (prune) renderer like the reference's GPU path
    (:170); otherwise the FlatRenderer (:160-168) -- both run on the B200 here."""

    def __init__(self, STLOutput=None, Resolution=0.0, UseGPU=True, Silent=False):
        self.STLOutput, self.Resolution, self.UseGPU, self.Silent = STLOutput, float(Resolution), UseGPU, Silent


def RenderShader3D(s, cfg):
    if not (cfg.Resolution > 0) or not np.isfinite(cfg.Resolution):
        raise ValueError("RenderConfig resolution must be positive, non-infinity")       # gsdfaux.go:64-66
    if cfg.STLOutput is None:
        raise ValueError("RenderShader3D requires output parameter in config")           # :71-73

    def log(elapsed, *args):
        if not cfg.Silent:
            print("[%s]" % ("%.3fms" % (elapsed * 1e3) if elapsed else "-"), *args)

    start = time.perf_counter()
    log(0, "using CUDA (sm_100a)")
    sdf = gleval.NewCUDASDF3(s)
    log(time.perf_counter() - start, "instantiating evaluation SDF took")
    t0 = time.perf_counter()
    renderer = glrender.NewOctreeRenderer(sdf, np.float32(cfg.Resolution), 1 << 15) if cfg.UseGPU else \
        glrender.NewFlatRenderer(sdf, np.float32(cfg.Resolution), 4096, 1)
    triangles = glrender.RenderAll(renderer)
    ev = renderer.Evaluations()
    if isinstance(renderer, glrender.Octree):
        omitted = 8 * renderer.TotalPruned()
        pct = 100.0 * omitted / max(ev + omitted, 1)
        log(time.perf_counter() - t0, "evaluated SDF", ev, "times and rendered", len(triangles), "triangles with", "%.2f" % pct,
            "percent evaluations omitted in octree pruning step with resolution", np.float32(cfg.Resolution))     # :219-224
    else:
        log(time.perf_counter() - t0, "evaluated SDF", ev, "times and rendered", len(triangles), "triangles with resolution",
            np.float32(cfg.Resolution))                                                                           # :225-226
    t0 = time.perf_counter()
    cfg.STLOutput.write(renderer.STLBytes())   # WriteBinarySTL, packed on the device, ONE write (stl.go:53 does one per triangle)
    log(time.perf_counter() - t0, "wrote STL")
    log(time.perf_counter() - start, "render done")
    return triangles
