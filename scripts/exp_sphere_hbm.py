"""BASELINE config 1's tree (single sphere) through gleval.SDF3.Evaluate on device-resident points: the one
HBM-bound case (16 B/eval: 12 read + 4 written). Prints achieved GB/s against MEASURED_PEAKS.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gsdf_b200 import gsdf, gleval
b = gsdf.Builder()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6482.7) if os.path.exists("MEASURED_PEAKS.json") else 6482.7
for name, s in (("sphere", b.NewSphere(1.0)), ("box", b.NewBox(1, 1.2, 1.5, 0.1)), ("union(sphere,box)", b.Union(b.NewSphere(1.0), b.Translate(b.NewBox(1, 1.2, 1.5, 0.1), 0.5, 0, 0)))):
    sdf = gleval.NewCUDASDF3(s)
    for n in (1 << 26,):
        pos = torch.rand(n, 3, device="cuda") * 4 - 2
        out = torch.empty(n, device="cuda")
        for _ in range(3):
            sdf.Evaluate(pos, out)
        torch.cuda.synchronize()
        ms = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sdf.Evaluate(pos, out); e1.record(); torch.cuda.synchronize()   # kernel runs on torch's current stream
            ms.append(e0.elapsed_time(e1))
        m = float(np.median(ms))
        if name == "sphere":
            ref = pos.double().norm(dim=1) - 1
            assert float((out.double() - ref).abs().max()) < 2e-6
        print(json.dumps(dict(tree=name, points=n, ms=m, Gevals_s=n / m / 1e6, GBs=16.0 * n / m / 1e6, frac_of_measured_hbm=16.0 * n / m / 1e6 / peak)))
