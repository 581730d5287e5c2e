"""Experiment: end-to-end (host buffers) time of one flange@400 render, single renderer vs Z-slab pipelines."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gsdf_b200 import gsdf, gleval, glrender, _lib
b = gsdf.Builder(); s = gsdf.scene(b, "npt-flange")
res = np.float32(s.Diagonal() / np.float32(400))
sdf = gleval.NewCUDASDF3(s)
R = glrender.NewOctreeRenderer(sdf, res, 1 << 15)
ntri = R.NumTriangles()
host = torch.empty((ntri + 8, 3, 3), dtype=torch.float32).pin_memory().numpy()
flat = b.flatten(s); blob, aux = flat["blob"], np.ascontiguousarray(flat["aux"]); auxp = aux.ctypes.data_as(C.POINTER(C.c_float))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, k=100):
    for _ in range(5): fn()
    ts = []
    for _ in range(k):
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e3, np.mean(ts) * 1e3
def single():
    _lib.check(_lib.lib.gsdf_program_update(sdf._h, blob, len(blob), auxp, aux.size))
    R.Rerun()
    n = _lib.lib.gsdf_mesh_read(R._h, C.c_void_p(host.ctypes.data), ntri + 8)
    assert n == ntri
ref = None
print("single: median %.3f ms mean %.3f ms" % timeit(single))
ref = host[:ntri].copy()
for ns in (2, 3, 4, 6):
    P = glrender.SlabPipeline(sdf, res, nslabs=ns)
    def pipe():
        _lib.check(_lib.lib.gsdf_program_update(sdf._h, blob, len(blob), auxp, aux.size))
        n = P.RenderToHost(host)
        assert n == ntri
    print("pipeline %d slabs: median %.3f ms mean %.3f ms" % ((ns,) + timeit(pipe)), "equal:", bool(np.array_equal(host[:ntri], ref)))
    P.Close()
# dual contour timing at flange resdiv 200 / 400
for rd in (100, 200, 400):
    r = np.float32(s.Diagonal() / np.float32(rd))
    d = glrender.DualContourRenderer()
    for placer in (glrender.DualContourNaive(), glrender.DualContourLeastSquares(Chiseled=True)):
        d.Reset(sdf, r, placer); d.Rerun()
        print("dual contour resdiv", rd, type(placer).__name__, d.Stats())
    d.Close()
