"""Mirror of the reference's driver gsdfaux.RenderShader3D (gsdfaux/gsdfaux.go:63-241) for the CUDA backend: pick the
evaluator, pick the renderer, RenderAll, log what the reference logs, WriteBinarySTL. It is the CALLER of the hot
path, kept thin; bench.py measures the same sequence."""
import time

import numpy as np

from . import gleval, glrender


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
