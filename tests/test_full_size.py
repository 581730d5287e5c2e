"""BASELINE.json's configurations at their STATED sizes, CUDA path against the oracle (-m gpu):

  config 2  npt-flange resdiv 400        whole lattice: cube-case indices, triangles, STL bytes, evaluation counts
  config 3  bolt resdiv 800              Z-slabs spread over the height (oracle side bounded), sha256 of triangle bytes
  config 4  knurled-cylinder resdiv 1600 Z-slabs (the lattice is 448 M corners: one slab of it is what a rank of the 8-GPU
                                         partition meshes), sha256 of triangle bytes and case indices
  config 5  text "Abc123~" at 8192 x 8192  bands of rows against the oracle, the whole image guarded == unguarded

The oracle evaluates only the corner planes of the slabs / the rows of the bands it compares (oracle.flat_eval_planes,
Tree.eval2), with the same absolute lattice positions as a whole-lattice sweep."""
import hashlib
import os

import numpy as np
import pytest

import gsdf_b200
from gsdf_b200 import gsdf, gleval, glrender
import fontfix

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_flange_resdiv400_whole_lattice(oracle, bld):
    """The headline configuration end to end: 281x281x85 corners, 423,852 triangles (README.md:116,130)."""
    s = gsdf.scene(bld, "npt-flange")
    res = np.float32(s.Diagonal() / np.float32(400))
    t = oracle.Tree.from_shader(s)
    lat = oracle.flat_lattice(*s.Bounds(), res)
    assert list(lat.n) == [280, 280, 84]
    grid, ev = oracle.flat_eval_grid(t, lat, nthreads=NT)
    assert ev == 6711685
    want, wcases = oracle.flat_march(lat, grid, want_cases=True)
    assert len(want) == 423852
    sdf = gleval.NewCUDASDF3(s)
    F = glrender.FlatRenderer(sdf, res, keep_cases=True, keep_grid=True)
    assert np.array_equal(bits(F.Grid()), bits(grid))                     # all 6,711,685 distances bit-equal
    assert np.array_equal(F.Cases(), wcases)
    assert np.array_equal(bits(F.AllTriangles()), bits(want))
    wstl = oracle.stl_write(want)
    assert F.STLBytes() == wstl
    for prune in (None, "literal"):
        R = glrender.Octree(sdf, res, keep_cases=True, prune=prune)
        mask, kept, centres = oracle.octree_prune_plan(t, lat, R.Plan())
        wp, wpc = oracle.flat_march(lat, grid, want_cases=True, blockmask=mask)
        assert len(wp) == 423852                                          # the prune loses nothing here
        assert np.array_equal(R.Cases(), wpc)
        assert np.array_equal(bits(R.AllTriangles()), bits(want))
        assert R.STLBytes() == wstl
        assert R.TotalPruned() == (mask.size - kept) * (1 << (R.Plan()[-1][0] - 1)) ** 3
        assert centres < R.Evaluations() < ev // 3
        R.Close()
    F.Close()


def slab_check(oracle, t, lat, sdf, res, cz0, cz1, prune_plan):
    planes = oracle.flat_eval_planes(t, lat, cz0, cz1 + 1, nthreads=NT)
    mask = None
    if prune_plan:
        mask, _, _ = oracle.octree_prune_plan(t, lat, prune_plan)
    want, wcases = oracle.flat_march_planes(lat, planes, cz0, cz1, blockmask=mask, want_cases=True)
    return planes, want, wcases


@pytest.mark.parametrize("scene,resdiv,lattice,slabs", [
    ("bolt", 800, None, [(0, 12), (150, 166), (301, 313), (514, 526)]),
    ("knurled-cylinder", 1600, [563, 563, 1407], [(0, 8), (700, 708), (1399, 1407)]),
])
def test_large_configs_by_z_slab(oracle, bld, scene, resdiv, lattice, slabs):
    s = gsdf.scene(bld, scene)
    res = np.float32(s.Diagonal() / np.float32(resdiv))
    t = oracle.Tree.from_shader(s)
    lat = oracle.flat_lattice(*s.Bounds(), res)
    if lattice:
        assert list(lat.n) == lattice
    nz = lat.n[2]
    sdf = gleval.NewCUDASDF3(s)
    for cz0, cz1 in slabs:
        cz1 = min(cz1, nz)
        F = glrender.FlatRenderer(sdf, res, cz_range=(cz0, cz1), keep_cases=True, keep_grid=True)
        planes, want, wcases = slab_check(oracle, t, lat, sdf, res, cz0, cz1, None)
        assert np.array_equal(bits(F.Grid()), bits(planes)), (scene, cz0)
        assert sha(F.Cases()) == sha(wcases), (scene, cz0)
        got = F.AllTriangles()
        assert len(got) == len(want) and sha(got) == sha(want), (scene, cz0, len(got), len(want))
        F.Close()
        # the Octree renderer (default plan; on these lattices it has a coarse level in front of level 3) against the
        # oracle's restatement of the same plan -- and against the dense slab, which it must reproduce
        R = glrender.Octree(sdf, res, cz_range=(cz0, cz1), keep_cases=True)
        _, wantp, wpc = slab_check(oracle, t, lat, sdf, res, cz0, cz1, R.Plan())
        gotp = R.AllTriangles()
        assert sha(R.Cases()) == sha(wpc), (scene, cz0)
        assert len(gotp) == len(wantp) and sha(gotp) == sha(wantp), (scene, cz0)
        assert len(gotp) == len(want), (scene, cz0, "default prune plan lost triangles")
        R.Close()


def test_text_8192_bands_and_whole_image(oracle, bld, monkeypatch):
    """Config 5: TextLine("Abc123~") on 8192 x 8192 pixels. Bands of rows (top, bottom, through the glyph bodies) are
    compared with the oracle pixel by pixel; the whole image must be identical with and without the box guards (two
    different instruction streams over the same 67 M pixels)."""
    W = H = 8192
    s = fontfix.text_scene(bld)
    t = oracle.Tree.from_shader(s)
    mn, mx = s.Bounds()
    sdf = gleval.NewCUDASDF2(s)
    img = glrender.ImageEvaluateSDF2(sdf, W, H)
    f = np.float32
    dx = f((f(mx[0]) - f(mn[0])) / f(W)); dy = f((f(mx[1]) - f(mn[1])) / f(H))   # image.go:85-87
    xmin = f(f(mn[0]) + f(dx / f(2)))
    xs = (np.arange(W, dtype=np.float32) * dx + xmin).astype(np.float32)
    for j0 in (0, 1000, 2731, 4096, 5555, 7000, H - 16):
        rows = np.arange(j0, j0 + 16)
        ys = (f(mx[1]) - rows.astype(np.float32) * dy).astype(np.float32)                 # image.go:92
        pos = np.stack([np.tile(xs, len(rows)), np.repeat(ys, W)], 1).astype(np.float32)
        want = t.eval2(pos).reshape(len(rows), W)
        assert np.array_equal(bits(img[j0:j0 + 16]), bits(want)), j0
    digest = sha(img)
    del img
    monkeypatch.setenv("GSDF_NO_GUARDS", "1")
    plain = gleval.NewCUDASDF2(s)
    assert sha(glrender.ImageEvaluateSDF2(plain, W, H)) == digest
