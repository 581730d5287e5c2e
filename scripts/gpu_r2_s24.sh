#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/ab_flange.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from gsdf_b200 import gsdf, gleval, glrender, _lib
from oracle import oracle as O
b = gsdf.Builder()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
s = gsdf.scene(b, "npt-flange")
sdf = gleval.NewCUDASDF3(s)
res = np.float32(s.Diagonal() / np.float32(400))
R = glrender.Octree(sdf, res)
ts = []
for i in range(12):
    flush.fill_(1); torch.cuda.synchronize()
    R.Rerun(); ts.append(R.Timings())
t = {k: float(np.median([x[k] for x in ts[3:]])) for k in ts[0]}
tris = R.AllTriangles()
import hashlib
print(os.environ.get("GSDF_SPECIAL"), os.environ.get("GSDF_SPECIAL_CTA"), os.path.basename(_lib.LIB_PATH), "tris", len(tris), hashlib.sha256(tris.tobytes()).hexdigest()[:16],
      "prune %.4f eval %.4f classify %.4f emit %.4f total %.4f" % (t["prune_ms"], t["eval_ms"], t["classify_ms"], t["emit_ms"], t["total_ms"]))
PY
: > gpurun_out/ab_special.txt
L=$PWD/gsdf_b200/libgsdfb200_special.so
python /tmp/ab_flange.py >> gpurun_out/ab_special.txt 2>&1
GSDF_B200_LIB=$L python /tmp/ab_flange.py >> gpurun_out/ab_special.txt 2>&1
for c in 128 192 256 384; do GSDF_B200_LIB=$L GSDF_SPECIAL=1 GSDF_SPECIAL_CTA=$c python /tmp/ab_flange.py >> gpurun_out/ab_special.txt 2>&1; done
cat gpurun_out/ab_special.txt
