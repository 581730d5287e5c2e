"""Child process of test_multi_device.py::test_device_driven_read_back_matches_the_single_renderer: with GSDF_MULTI_COPYK=1
(read once per process) the multi-slab driver enqueues every slab's read-back up front as a kernel that reads the triangle
counts itself (mesher.cu k_copy_out). First renders (buffers grow, the emit pass runs twice: classic copy) and steady-state
renders (copy kernels), 1 to 7 slabs, must fill the pinned destination exactly as the single renderer's AllTriangles."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gsdf_b200 import gsdf, gleval, glrender  # noqa: E402

bld = gsdf.Builder()
ok = True
for scene, resdiv in [("npt-flange", 150), ("knurled-cylinder", 130), ("bolt", 120)]:
    s = gsdf.scene(bld, scene)
    res = np.float32(s.Diagonal() / np.float32(resdiv))
    sdf = gleval.NewCUDASDF3(s)
    want = glrender.Octree(sdf, res).AllTriangles()
    for spd in (1, 2, 3, 7):
        M = glrender.MultiRenderer(s, res, devices=[0], slabs_per_device=spd)
        host = glrender.pinned_empty((len(want) + 8, 3, 3))
        for run in range(4):
            host[:] = np.nan
            n = M.RenderToHost(host)
            same = n == len(want) and np.array_equal(host[:n].view(np.uint32), want.view(np.uint32)) and bool(np.isnan(host[n:]).all())
            ok = ok and same
            if not same:
                print("MISMATCH", scene, "slabs", spd, "run", run, n, len(want))
            if run == 1:
                M.Rebalance(2)  # new cuts: fresh slab handles, first renders again
        pageable = np.full((len(want) + 8, 3, 3), np.nan, np.float32)
        n = M.RenderToHost(pageable)
        same = n == len(want) and np.array_equal(pageable[:n].view(np.uint32), want.view(np.uint32))
        ok = ok and same
        if not same:
            print("MISMATCH pageable", scene, "slabs", spd)
        M.Close()
print("MULTI COPYK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
