"""Phase timing of a 2-slab pipeline (where does the time go?)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gsdf_b200 import gsdf, gleval, glrender, _lib
from gsdf_b200._lib import lib, check
b = gsdf.Builder(); s = gsdf.scene(b, "npt-flange")
res = np.float32(s.Diagonal() / np.float32(400))
sdf = gleval.NewCUDASDF3(s)
P = glrender.SlabPipeline(sdf, res, nslabs=2)
ntri = P.NumTriangles()
host = torch.empty((ntri + 8, 3, 3), dtype=torch.float32).pin_memory().numpy()
flat = host.reshape(-1)
for _ in range(5): P.RenderToHost(host)
acc = np.zeros(8)
K = 50
for _ in range(K):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    got = 0
    for p in P.parts:
        check(lib.gsdf_mesh_rerun(p._h)); t.append(time.perf_counter())
        n = check(lib.gsdf_mesh_read_async(p._h, C.c_void_p(flat[9 * got:].ctypes.data), ntri + 8 - got)); t.append(time.perf_counter())
        got += n
    for p in P.parts:
        check(lib.gsdf_mesh_wait(p._h)); t.append(time.perf_counter())
    acc[:len(t) - 1] += np.diff(t)
print("phases us: rerunA %.0f readasyncA %.0f rerunB %.0f readasyncB %.0f waitA %.0f waitB %.0f" % tuple(acc[:6] / K * 1e6))
print("device ms per part:", [p.Timings()["total_ms"] for p in P.parts], "tris", [p.NumTriangles() for p in P.parts])
