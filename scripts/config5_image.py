"""BASELINE config 5 at full size: forge/textsdf TextLine("Abc123~") (RelativeGlyphTolerance 0.001,
examples/image-text/text.go:24-34) evaluated at 8192 x 8192 pixel centres with ImageRendererSDF2's positions
(glrender/image.go:76-105), distances and fused RGBA rendering. Device time by CUDA events, end-to-end wall time
including the 268 MB device->host copy, and a bounded CPU sample of the oracle (rows of the same image) beside it.
Run under gpurun; appends one JSON line to gpurun_out/config5.jsonl."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import fontfix
from gsdf_b200 import gsdf, gleval, glrender, gsdfaux, _lib
from gsdf_b200._lib import check, lib
from oracle import oracle as O

W = H = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
b = gsdf.Builder()
s = fontfix.text_scene(b)
sdf = gleval.NewCUDASDF2(s)
f = b.flatten(s)
nedges = sum(n.aux_cnt // 2 for n in b.tree_nodes() if n.kind == gsdf.K["POLY2D"])
mn, mx = s.Bounds()
a = (C.c_float * 2)(float(mn[0]), float(mn[1]))
bb = (C.c_float * 2)(float(mx[0]), float(mx[1]))
edge = np.float32(mx[1] - mn[1]) / np.float32(1000)
conv = gsdfaux.ColorConversionLinearGradient(edge, gsdfaux.Black, gsdfaux.White)

# parity at a size the oracle finishes quickly
small = glrender.ImageEvaluateSDF2(sdf, 512, 128)
want = O.Tree.from_shader(s).image_eval2(mn, mx, 512, 128)
parity = bool(np.array_equal(small.view(np.uint32), want.view(np.uint32)))

st = torch.cuda.Stream()
d_out = torch.empty(W * H, dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
with torch.cuda.stream(st):
    for name, call in (("dist", lambda: check(lib.gsdf_image_eval2_device(sdf._h, a, bb, W, H, C.c_void_p(d_out.data_ptr()), C.c_void_p(st.cuda_stream)))),
                       ("rgba", lambda: check(lib.gsdf_image_render2_device(sdf._h, a, bb, W, H, C.byref(conv), C.c_void_p(d_out.data_ptr()), C.c_void_p(st.cuda_stream))))):
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); call(); e1.record(st)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        res[name + "_device_ms"] = float(np.median(ms))
# end to end (host buffers): ImageEvaluateSDF2 and the fused RGBA render
host = np.empty((H, W, 4), np.uint8)
r = glrender.NewImageRendererSDF2(max(4096, W), conv)
r.Render(sdf, host)
t0 = time.perf_counter(); r.Render(sdf, host); t1 = time.perf_counter()
res["rgba_e2e_ms"] = (t1 - t0) * 1e3
# bounded CPU sample: 8 rows of the same image through the oracle, one thread
rows = 8
tree = O.Tree.from_shader(s)
dy = np.float32(mx[1] - mn[1]) / np.float32(H)
t0 = time.perf_counter()
tree.image_eval2(mn, np.array([mx[0], mn[1] + dy * rows], np.float32), W, rows)
cpu_s = time.perf_counter() - t0
rec = dict(config="textsdf 'Abc123~' %dx%d" % (W, H), program={k: f[k] for k in ("ninstr", "nchunks", "dstack", "pstack")},
           polygon_edges=nedges, aux_floats=int(f["aux"].size), parity_512x128_bit_equal=parity, pixels=W * H, **res,
           dist_Gpix_s=W * H / res["dist_device_ms"] / 1e6, rgba_Gpix_s=W * H / res["rgba_device_ms"] / 1e6,
           hbm_write_GBs=W * H * 4 / res["dist_device_ms"] / 1e6,
           cpu_port_1thread_Mpix_s=W * rows / cpu_s / 1e6, cpu_sample="%d rows of %d pixels, %.1f s" % (rows, W, cpu_s),
           inside_fraction=float((host[..., 0] < 128).mean()))
print(json.dumps(rec))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/config5.jsonl", "a") as fp:
    fp.write(json.dumps(rec) + "\n")
