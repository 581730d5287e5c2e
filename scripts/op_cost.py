"""Per-node cost of the interpreter: dense-lattice evaluation rate of small programs (CUDA events, 16.7 M points)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gsdf_b200 import gsdf, gleval, glrender, _lib
import shapes
b = gsdf.Builder()
T = gsdf.threads
items = [("sphere", b.NewSphere(1)), ("box", b.NewBox(1, .5, .8, .1)), ("cylinder", b.NewCylinder(1, .5, 0)),
         ("cylinder_round", b.NewCylinder(1, .5, .1)), ("torus", b.NewTorus(1, .3)),
         ("translate+sphere", b.Translate(b.NewSphere(1), .1, .2, .3)),
         ("union2 spheres", b.Union(b.NewSphere(1), b.Translate(b.NewSphere(1), 1, 0, 0))),
         ("smoothunion2", b.SmoothUnion(.1, b.NewSphere(1), b.Translate(b.NewSphere(1), 1, 0, 0))),
         ("extrude poly12", b.Extrude(T.Thread(b, T.NPT(0.5)), 1)),
         ("extrude poly18", b.Extrude(T.Thread(b, T.ISO(3, .5, True)), 1)),
         ("extrude hexagon6", b.Extrude(b.NewPolygon(shapes.nagon(6, 1)), 1)),
         ("screw npt", T.Screw(b, .43, T.NPT(0.5))),
         ("twist box", b.Twist(b.NewBox(1, 1, 2, 0), .5)), ("circarray box", b.CircularArray(b.Translate(b.NewBox(.3, .3, 1, 0), 1, 0, 0), 8, 8)),
         ("rotate box", b.Rotate(b.NewBox(1, .5, .8, 0), .7, (1, 2, 3))),
         ("npt-flange", gsdf.scene(b, "npt-flange")), ("bolt", gsdf.scene(b, "bolt")), ("knurled", gsdf.scene(b, "knurled-cylinder"))]
N = 255
for name, s in items:
    sdf = gleval.NewCUDASDF3(s)
    mn, mx = s.Bounds()
    res = np.float32(float((mx - mn).max()) * 1.01 / N)
    lat = glrender.lattice_from_bounds(mn, mx, res)
    n = (lat.n[0] + 1) * (lat.n[1] + 1) * (lat.n[2] + 1)
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    ts = torch.cuda.Stream()
    st = ts.cuda_stream
    def run():
        _lib.check(_lib.lib.gsdf_grid_eval_device(sdf._h, C.byref(lat), 0, lat.n[2] + 1, C.c_void_p(out.data_ptr()), C.c_void_p(st)))
    for _ in range(2): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(ts)
    for _ in range(5): run()
    e1.record(ts); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    f = b.flatten(s)
    print("%-18s instr=%3d  %8.3f ms for %9d evals  -> %7.1f Geval/s  %6.3f ns/eval  (%.0f GB/s written)" % (name, f["ninstr"], ms, n, n / ms / 1e6, ms * 1e6 / n, 4 * n / ms / 1e6))
