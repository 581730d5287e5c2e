"""A/B timing of kernel variants: GSDF_B200_LIB=<so> python scripts/ab_eval.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gsdf_b200 import gsdf, gleval, glrender, _lib
b = gsdf.Builder()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, resdiv in [("npt-flange", 400), ("bolt", 400), ("knurled-cylinder", 500)]:
    s = gsdf.scene(b, name)
    sdf = gleval.NewCUDASDF3(s)
    if os.environ.get("GSDF_AB_SPECIAL"):
        print("specialised:", sdf.Specialize(), flush=True)
    res = np.float32(s.Diagonal() / np.float32(resdiv))
    for cls in (glrender.Octree, glrender.FlatRenderer):
        R = cls(sdf, res, stage_timing=os.environ.get("GSDF_AB_GRAPH") is None)
        ts = []
        for i in range(8):
            flush.fill_(1); torch.cuda.synchronize()
            R.Rerun(); ts.append(R.Timings())
        t = {k: float(np.median([x[k] for x in ts[2:]])) for k in ts[0]}
        print("%-28s %-18s %-12s evals=%9d tris=%8d  prune %.3f eval %.3f classify %.3f emit %.3f total %.3f ms  -> %.1f Geval/s executed" % (
            os.path.basename(_lib.LIB_PATH), name, cls.__name__, R.Evaluations(), R.NumTriangles(), t["prune_ms"], t["eval_ms"], t["classify_ms"], t["emit_ms"], t["total_ms"],
            R.Evaluations() / max(t["eval_ms"], 1e-9) / 1e6))
        R.Close()
