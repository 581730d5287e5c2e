#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference is pure Go and cannot run in this image, and it holds no golden distance arrays of its own
(SURVEY.md section 4), so these fixtures are ORACLE-generated regression pins: they freeze the oracle's bits after it
was pinned against the reference's known answers (41072 / 423,852 triangles, 6,711,685 evaluations, thread signs).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import shapes  # noqa: E402
from gsdf_b200 import gsdf  # noqa: E402
from oracle import oracle as O  # noqa: E402


def distances():
    b = gsdf.Builder()
    out = {}
    rng = np.random.default_rng(20261017)
    for name, s in shapes.all3d(b) + shapes.all2d(b):
        mn, mx = s.Bounds()
        d = len(mn)
        pos = (mn - 0.2 * (mx - mn) + rng.random((96, d), dtype=np.float32) * 1.4 * (mx - mn)).astype(np.float32)
        t = O.Tree.from_shader(s)
        out[name + ".pos"] = pos
        out[name + ".dist"] = t.eval3(pos) if d == 3 else t.eval2(pos)
    return out


def meshes():
    b = gsdf.Builder()
    out = {}
    cases = [("sphere", b.NewSphere(1.0), np.float32(1.0 / 33)),
             ("npt-flange", gsdf.scene(b, "npt-flange"), None),
             ("bolt", gsdf.scene(b, "bolt"), None),
             ("knurled-cylinder", gsdf.scene(b, "knurled-cylinder"), None)]
    for name, s, res in cases:
        if res is None:
            res = np.float32(s.Diagonal() / np.float32(120))
        t = O.Tree.from_shader(s)
        lat = O.flat_lattice(*s.Bounds(), res)
        grid, _ = O.flat_eval_grid(t, lat, nthreads=os.cpu_count() or 1)
        mask, kept = O.octree_prune_mask(t, lat)
        tris, cases_ = O.flat_march(lat, grid, want_cases=True)
        ptris, _ = O.flat_march(lat, grid, blockmask=mask)
        out[name + ".res"] = np.float32(res)
        out[name + ".n"] = np.array(list(lat.n), np.int32)
        out[name + ".ntri"] = np.int64(len(tris))
        out[name + ".ntri_pruned"] = np.int64(len(ptris))
        out[name + ".kept_blocks"] = np.int64(kept)
        out[name + ".tri_sha"] = np.frombuffer(hashlib.sha256(tris.tobytes()).digest(), np.uint8)
        out[name + ".case_sha"] = np.frombuffer(hashlib.sha256(cases_.tobytes()).digest(), np.uint8)
        out[name + ".stl_sha"] = np.frombuffer(hashlib.sha256(O.stl_write(tris)).digest(), np.uint8)
        out[name + ".first_tris"] = tris[:8].copy()
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "distances.npz"), **distances())
    np.savez_compressed(os.path.join(HERE, "meshes.npz"), **meshes())
    print("wrote", os.listdir(HERE))
