"""Child process of test_gpu_graph.py::test_count_kernel_tile_loop_matches_oracle: with GSDF_COUNT_GRID capping the
classification kernel's grid, every CTA walks many tiles, so the tile loop (mbarrier phases, stencil reuse, active
and pruned tiles in every order) runs on lattices small enough for the oracle. Cube-case indices and triangles must be
the oracle's bit for bit, eager and as a CUDA-graph replay."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gsdf_b200 import gsdf, gleval, glrender  # noqa: E402
from oracle import oracle as O  # noqa: E402

O.build()
bld = gsdf.Builder()
ok = True
for scene, resdiv in [("sphere", 70), ("npt-flange", 150), ("bolt", 120), ("knurled-cylinder", 130)]:
    s = bld.NewSphere(1.0) if scene == "sphere" else gsdf.scene(bld, scene)
    res = np.float32(s.Diagonal() / np.float32(resdiv))
    t = O.Tree.from_shader(s)
    lat = O.flat_lattice(*s.Bounds(), res)
    grid, _ = O.flat_eval_grid(t, lat, nthreads=os.cpu_count() or 1)
    sdf = gleval.NewCUDASDF3(s)
    if os.environ.get("GSDF_CHILD_SPECIALIZE"):  # the run-time compiled kernels (two corners per thread exists only there)
        assert sdf.Specialize(), "run-time compilation unavailable"
    for prune in (True, False):
        cases = os.environ.get("GSDF_CHILD_NO_CASES") is None  # (parity mode launches the count pass without a programmatic edge)
        R = (glrender.Octree if prune else glrender.FlatRenderer)(sdf, res, keep_cases=cases)
        mask = O.octree_prune_plan(t, lat, R.Plan())[0] if prune else None
        wt, wc = O.flat_march(lat, grid, want_cases=True, blockmask=mask)
        for run in range(4):  # eager, graph capture, graph replays
            if run:
                R.Rerun()
            tris = R.AllTriangles()
            same = (not cases or int((R.Cases() != wc).sum()) == 0) and len(tris) == len(wt) and np.array_equal(tris.view(np.uint32), wt.view(np.uint32))
            ok = ok and same
            if not same:
                print("MISMATCH", scene, "prune" if prune else "flat", "run", run, len(tris), len(wt))
        R.Close()
print("COUNT PIPELINE", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
