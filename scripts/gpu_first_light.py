"""First-light check on a GPU box: CUDA path vs CPU oracle on sphere / flange / knurled (eval + mesh)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gsdf_b200
from gsdf_b200 import gsdf, gleval, glrender
from oracle import oracle as O

print(gsdf_b200.version(), "devices", gsdf_b200.device_count())
b = gsdf.Builder()
rng = np.random.default_rng(1)

def cmp_eval(name, s, n=200000):
    mn, mx = s.Bounds()
    pos = (rng.random((n, 3), dtype=np.float32) * (mx - mn) * 1.2 + (mn - 0.1 * (mx - mn))).astype(np.float32)
    t = O.Tree.from_shader(s)
    want = t.eval3(pos)
    sdf = gleval.NewCUDASDF3(s)
    got = np.empty(n, np.float32)
    sdf.Evaluate(pos, got)
    bad = np.flatnonzero(got.view(np.uint32) != want.view(np.uint32))
    err = np.abs(got - want) / np.maximum(1, np.abs(want))
    print("%-18s n=%d bit-mismatch=%d max-rel-err=%.3g" % (name, n, bad.size, err.max()))
    if bad.size:
        i = bad[0]; print("   first mismatch", pos[i], got[i], want[i])
    return sdf

def cmp_mesh(name, s, resdiv, prune):
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(s.Diagonal() / np.float32(resdiv))
    t0 = time.time()
    R = (glrender.Octree if prune else glrender.FlatRenderer)(sdf, res, keep_cases=True, keep_grid=not prune)
    t1 = time.time()
    tris = R.AllTriangles()
    t2 = time.time()
    t = O.Tree.from_shader(s)
    lat = O.flat_lattice(*s.Bounds(), res)
    assert list(lat.n) == list(R.lat.n), (list(lat.n), list(R.lat.n))
    grid, ev = O.flat_eval_grid(t, lat, nthreads=os.cpu_count())
    mask = O.octree_prune_mask(t, lat)[0] if prune else None
    wt, wc = O.flat_march(lat, grid, want_cases=True, blockmask=mask)
    cases = R.Cases()
    print("%-18s resdiv=%d prune=%d lattice=%s tris gpu=%d oracle=%d cases-mismatch=%d tri-bits-equal=%s evals=%d pruned=%d begin=%.1fms read=%.1fms %s" % (
        name, resdiv, prune, list(lat.n), len(tris), len(wt), int((cases != wc).sum()),
        len(tris) == len(wt) and bool((tris.view(np.uint32) == wt.view(np.uint32)).all()), R.Evaluations(), R.TotalPruned(),
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, R.Timings()))
    if not prune:
        g = R.Grid()
        print("   grid bit-mismatch:", int((g.view(np.uint32) != grid.view(np.uint32)).sum()), "of", g.size)
    for _ in range(3):
        R.Rerun()
    print("   rerun timings", R.Timings())
    stl = R.STLBytes()
    want_stl = O.stl_write(wt)
    print("   stl bytes", len(stl), "equal to oracle:", stl == want_stl)

sph = b.NewSphere(1.0)
cmp_eval("sphere", sph)
fl = gsdf.scene(b, "npt-flange")
cmp_eval("npt-flange", fl)
bo = gsdf.scene(b, "bolt")
cmp_eval("bolt", bo)
kn = gsdf.scene(b, "knurled-cylinder")
cmp_eval("knurled", kn)
cmp_mesh("sphere", sph, 115, False)
cmp_mesh("sphere", sph, 115, True)
cmp_mesh("npt-flange", fl, 400, False)
cmp_mesh("npt-flange", fl, 400, True)
cmp_mesh("bolt", bo, 200, True)
cmp_mesh("knurled", kn, 300, True)
