"""e2e of the multi-slab pipeline on one GPU for the slab counts in argv (default 1 2 3 4), timeline of the last render.
Environment knobs (GSDF_PDL, GSDF_SCAN_FUSED, ...) are read by the library once per process: run one process per setting."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gsdf_b200 import gsdf, gleval, glrender
b = gsdf.Builder()
s = gsdf.scene(b, "npt-flange")
res = np.float32(s.Diagonal() / np.float32(400))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
f = b.flatten(s); blob, aux = f["blob"], np.ascontiguousarray(f["aux"])
ntri = 423852
host = glrender.pinned_empty((ntri + 8, 3, 3))
tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("GSDF_") and k != "GSDF_B200_LIB")
for spd in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]:
    M = glrender.MultiRenderer(s, res, devices=[0], slabs_per_device=spd)
    if os.environ.get("GSDF_AB_SPECIAL"):
        M.Specialize()
    for reb in (0, 1):
        if reb:
            if spd == 1: continue
            M.Rebalance(2)
        ts = []
        for i in range(40):
            flush.fill_(1); torch.cuda.synchronize()
            t0 = time.perf_counter(); M.UpdateBlob(blob, aux); n = M.RenderToHost(host); ts.append(time.perf_counter() - t0)
            assert n == ntri
        tl = M.Timeline()
        print("[%s] slabs=%2d reb=%d  median %.4f min %.4f ms  device %.4f ms  cuts %s | enq %.0f | " % (tag, spd, reb, np.median(ts[5:]) * 1e3, min(ts) * 1e3, M.DeviceMs(), M.Slabs()[0], tl["enqueued"]) +
              " ".join("(%.0f,%.0f)" % x for x in tl["slabs"]) + " | delivered %.0f" % tl["delivered"], flush=True)
    M.Close()
