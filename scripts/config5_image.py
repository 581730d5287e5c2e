"""BASELINE config 5's evaluator path at full size: a polygon-heavy 2-D tree (7 'glyphs' = 9 polygons of 20..56 vertices,
union / difference / translate2D, like forge/textsdf builds for "Abc123~") evaluated at 8192 x 8192 pixel centres with
ImageRendererSDF2's positions (glrender/image.go:76-105). Font parsing is out of scope; the polygons are synthetic."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gsdf_b200 import gsdf, gleval, glrender
from oracle import oracle as O

def blob(n, r, seed):
    rng = np.random.default_rng(seed)
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    rad = r * (1 + 0.25 * np.sin(3 * ang + seed) + 0.05 * rng.standard_normal(n))
    return np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1).astype(np.float32)

b = gsdf.Builder()
glyphs = []
for i, nv in enumerate([22, 33, 30, 20, 36, 56, 34]):   # control-point counts of "Abc123~" in ISO-3098
    g = b.NewPolygon(blob(nv, 0.4, i))
    if i < 2:                                             # 'A' and 'b' have a second contour (a hole)
        g = b.Difference2D(g, b.NewPolygon(blob(12, 0.12, 10 + i)))
    glyphs.append(b.Translate2D(g, 0.9 * i, 0))
s = b.Union2D(*glyphs)
sdf = gleval.NewCUDASDF2(s)
f = b.flatten(s)
print("program:", {k: f[k] for k in ("ninstr", "nchunks", "dstack", "pstack")}, "aux floats", f["aux"].size)
for W in (1024, 8192):
    H = W
    glrender.ImageEvaluateSDF2(sdf, W, H)
    t0 = time.perf_counter(); img = glrender.ImageEvaluateSDF2(sdf, W, H); t1 = time.perf_counter()
    print("image %dx%d: %.2f ms wall incl. %.0f MB D2H -> %.2f G pixels/s" % (W, H, (t1 - t0) * 1e3, img.nbytes / 1e6, W * H / (t1 - t0) / 1e9))
    if W == 1024:
        want = O.Tree.from_shader(s).image_eval2(*s.Bounds(), W, H)
        print("   bit-equal to oracle:", bool(np.array_equal(img.view(np.uint32), want.view(np.uint32))))
# device-resident rate through Evaluate on CUDA tensors
import torch
n = 8192 * 8192
mn, mx = s.Bounds()
xs = torch.rand(n, 2, device="cuda") * torch.tensor(mx - mn, device="cuda") + torch.tensor(mn, device="cuda")
out = torch.empty(n, device="cuda")
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    sdf.Evaluate(xs, out); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); sdf.Evaluate(xs, out); e1.record(st); torch.cuda.synchronize()
print("Evaluate (device resident) %d points: %.2f ms -> %.2f G evals/s" % (n, e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e6))
