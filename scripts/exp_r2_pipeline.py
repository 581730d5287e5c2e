"""Round-2 experiments on one GPU: (1) e2e of the multi-slab pipeline for several slab counts, (2) per-stage device times
(in-graph stamps) of one Z-slab as it gets thinner: the latency floor of the render chain."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gsdf_b200 import gsdf, gleval, glrender
b = gsdf.Builder()
s = gsdf.scene(b, "npt-flange")
res = np.float32(s.Diagonal() / np.float32(400))
sdf = gleval.NewCUDASDF3(s)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
f = b.flatten(s); blob, aux = f["blob"], np.ascontiguousarray(f["aux"])
ntri = 423852
host = glrender.pinned_empty((ntri + 8, 3, 3))
for spd in (1, 2, 3, 4, 6, 8, 12):
    M = glrender.MultiRenderer(s, res, devices=[0], slabs_per_device=spd)
    for reb in (0, 1):
        if reb:
            M.Rebalance(2)
        ts = []
        for i in range(40):
            flush.fill_(1); torch.cuda.synchronize()
            t0 = time.perf_counter(); M.UpdateBlob(blob, aux); n = M.RenderToHost(host); ts.append(time.perf_counter() - t0)
            assert n == ntri
        print("e2e slabs=%2d rebalanced=%d  median %.4f ms  min %.4f ms  device %.4f ms  cuts %s" % (spd, reb, np.median(ts[5:]) * 1e3, min(ts) * 1e3, M.DeviceMs(), M.Slabs()[0]))
        tl = M.Timeline()
        print("     timeline us: enqueued %.0f | " % tl["enqueued"] + " ".join("(%.0f,%.0f)" % x for x in tl["slabs"]) + " | delivered %.0f" % tl["delivered"])
    M.Close()
nz = 84
for layers in (84, 42, 20, 12, 8, 4):
    R = glrender.Octree(sdf, res, cz_range=(40 - min(40, layers // 2), min(nz, 40 - min(40, layers // 2) + layers)))
    tt = []
    for i in range(30):
        flush.fill_(1); torch.cuda.synchronize()
        R.Rerun(); tt.append(R.Timings())
    m = {k: float(np.median([x[k] for x in tt[5:]])) for k in tt[0]}
    print("slab of %2d layers: evals %8d tris %7d  prune %.4f eval %.4f classify+scan %.4f emit %.4f total %.4f ms" % (layers, R.Evaluations(), R.NumTriangles(), m["prune_ms"], m["eval_ms"], m["classify_ms"], m["emit_ms"], m["total_ms"]))
    R.Close()
