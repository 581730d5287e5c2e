"""Oracle vs the committed golden fixtures (tests/golden, made by make_golden.py): freezes the oracle's bits."""
import hashlib
import os

import numpy as np

import shapes
from gsdf_b200 import gsdf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_golden_distances(oracle, bld):
    g = np.load(os.path.join(GOLD, "distances.npz"))
    items = shapes.all3d(bld) + shapes.all2d(bld)
    assert len(items) >= 60
    for name, s in items:
        pos, want = g[name + ".pos"], g[name + ".dist"]
        t = oracle.Tree.from_shader(s)
        got = t.eval2(pos) if s.is2d else t.eval3(pos)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name


def test_oracle_reproduces_golden_meshes(oracle, bld):
    g = np.load(os.path.join(GOLD, "meshes.npz"))
    for name, s in [("sphere", bld.NewSphere(1.0)), ("bolt", gsdf.scene(bld, "bolt"))]:
        res = np.float32(g[name + ".res"])
        t = oracle.Tree.from_shader(s)
        lat = oracle.flat_lattice(*s.Bounds(), res)
        assert list(lat.n) == list(g[name + ".n"])
        grid, _ = oracle.flat_eval_grid(t, lat, nthreads=os.cpu_count() or 1)
        tris, cases = oracle.flat_march(lat, grid, want_cases=True)
        assert len(tris) == int(g[name + ".ntri"])
        assert hashlib.sha256(tris.tobytes()).digest() == g[name + ".tri_sha"].tobytes()
        assert hashlib.sha256(cases.tobytes()).digest() == g[name + ".case_sha"].tobytes()
        assert hashlib.sha256(oracle.stl_write(tris)).digest() == g[name + ".stl_sha"].tobytes()
    assert int(g["sphere.ntri"]) == 41072
