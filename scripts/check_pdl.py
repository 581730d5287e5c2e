"""Programmatic dependent launch check (on by default; GSDF_PDL=0 switches it off). Every renderer result (graph replay and eager launches, both
with the programmatic edges) must equal the stage-timed eager render, which always uses plain launches. Prints timings of
the graph replay so the same script serves as the A/B (GSDF_PDL=0 / 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gsdf_b200 import gsdf, gleval, glrender

b = gsdf.Builder()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ok = True
for name, resdiv in [("npt-flange", 400), ("npt-flange", 97), ("bolt", 300), ("knurled-cylinder", 300)]:
    s = gsdf.scene(b, name)
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(s.Diagonal() / np.float32(resdiv))
    for cls in (glrender.Octree, glrender.FlatRenderer):
        plain = cls(sdf, res, stage_timing=True)
        want = plain.AllTriangles().view(np.uint32)
        R = cls(sdf, res)
        ts = []
        for i in range(12):
            flush.fill_(1); torch.cuda.synchronize()
            R.Rerun()
            ts.append(R.Timings()["total_ms"])
            if i < 4 or i == 11:
                got = R.AllTriangles().view(np.uint32)
                same = got.shape == want.shape and np.array_equal(got, want) and (R.Evaluations(), R.TotalPruned()) == (plain.Evaluations(), plain.TotalPruned())
                ok = ok and same
                if not same: print("MISMATCH", name, resdiv, cls.__name__, "run", i)
        print("PDL=%s %-18s %4d %-12s tris=%8d graph step median %.4f ms min %.4f ms" % (os.environ.get("GSDF_PDL", "1"), name, resdiv, cls.__name__, R.NumTriangles(), float(np.median(ts[3:])), min(ts[3:])))
        R.Close(); plain.Close()
# the slab pipeline shares one stream between several meshers
s = gsdf.scene(b, "npt-flange")
sdf = gleval.NewCUDASDF3(s)
res = np.float32(s.Diagonal() / np.float32(200))
whole = glrender.Octree(sdf, res, stage_timing=True).AllTriangles().view(np.uint32)
P = glrender.SlabPipeline(sdf, res, 3)
dst = np.empty((len(whole) + 8, 3, 3), np.float32)
for i in range(5):
    dst[:] = 0
    n = P.RenderToHost(dst)
    got = dst[:n].reshape(-1).view(np.uint32)
    same = got.size == whole.size and np.array_equal(got, whole.reshape(-1))
    ok = ok and same
    if not same: print("MISMATCH slab pipeline run", i)
print("PDL CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
