"""gleval.PolygonGPU / Lines2DGPU / DisplaceMulti2D (gleval/gpu.go:169-446) and gleval.NormalsCentralDiff
(gleval/gleval.go:53-108) on the CUDA backend, against the oracle's evaluation of the equivalent node."""
import numpy as np
import pytest

import gsdf_b200
from gsdf_b200 import gsdf, gleval

pytestmark = pytest.mark.gpu


def grid2(mn, mx, n=64, pad=0.3):
    sz = mx - mn
    xs = np.linspace(mn[0] - pad * sz[0], mx[0] + pad * sz[0], n, dtype=np.float32)
    ys = np.linspace(mn[1] - pad * sz[1], mx[1] + pad * sz[1], n, dtype=np.float32)
    return np.ascontiguousarray(np.stack(np.meshgrid(xs, ys), -1).reshape(-1, 2))


def test_polygon_gpu(oracle, bld):
    verts = np.array([[0, 0], [2, 0], [2.5, 1.5], [1, 1], [0.2, 2]], np.float32)  # concave
    p = gleval.PolygonGPU(verts)
    with pytest.raises(gsdf_b200.GsdfError, match="invocation size"):
        p.Configure(gleval.ComputeConfig(InvocX=0))
    with pytest.raises(gsdf_b200.GsdfError, match="before Configure"):
        p.Evaluate(np.zeros((1, 2), np.float32), np.zeros(1, np.float32))
    p.Configure(gleval.ComputeConfig(InvocX=32))
    mn, mx = p.Bounds()
    assert np.array_equal(mn, verts.min(0)) and np.array_equal(mx, verts.max(0))
    pos = grid2(mn, mx)
    dist = np.empty(len(pos), np.float32)
    p.Evaluate(pos, dist)
    want = oracle.Tree.from_shader(bld.NewPolygon(verts)).eval2(pos)
    assert np.array_equal(dist.view(np.uint32), want.view(np.uint32))
    assert p.Evaluations() == len(pos)


def test_lines2d_gpu(oracle, bld):
    lines = np.array([[[0, 0], [1, 1]], [[2, 2], [3, 1]], [[-1, 0.5], [0.5, -1]]], np.float32)
    l = gleval.Lines2DGPU(lines, 0.25)
    l.Configure(gleval.ComputeConfig(InvocX=64))
    mn, mx = l.Bounds()
    assert np.allclose(mn, [-1.125, -1.125]) and np.allclose(mx, [3.125, 2.125])
    pos = grid2(mn, mx)
    dist = np.empty(len(pos), np.float32)
    l.Evaluate(pos, dist)
    want = oracle.Tree.from_shader(bld.NewLines2D(lines, 0.25)).eval2(pos)
    assert np.array_equal(dist.view(np.uint32), want.view(np.uint32))


def test_displace_multi2d(oracle, bld):
    disp = np.array([[0, 0], [3, 0.5], [-2, 2], [1, -3]], np.float32)
    d = gleval.DisplaceMulti2D(disp)
    d.Configure(lambda b: b.NewHexagon(0.7), gleval.ComputeConfig(InvocX=32))
    mn, mx = d.Bounds()
    pos = grid2(mn, mx)
    dist = np.empty(len(pos), np.float32)
    d.Evaluate(pos, dist)
    ref = bld.TranslateMulti2D(bld.NewHexagon(0.7), disp)
    want = oracle.Tree.from_shader(ref).eval2(pos)
    assert np.array_equal(dist.view(np.uint32), want.view(np.uint32))
    rmn, rmx = ref.Bounds()
    assert np.allclose(mn, rmn) and np.allclose(mx, rmx)


def test_normals_central_diff(oracle, bld):
    s = bld.SmoothUnion(0.2, bld.NewSphere(1.0), bld.Translate(bld.NewBox(1, 1, 1, 0.1), 0.8, 0.2, 0))
    sdf = gleval.NewCUDASDF3(s)
    rng = np.random.default_rng(3)
    pos = (rng.random((500, 3), dtype=np.float32) * 3 - 1.5).astype(np.float32)
    nrm = np.empty_like(pos)
    step = np.float32(1e-3)
    gleval.NormalsCentralDiff(sdf, pos, nrm, step)
    t = oracle.Tree.from_shader(s)
    h = step * np.float32(0.5)
    for dim in range(3):
        e = np.zeros(3, np.float32)
        e[dim] = h
        want = t.eval3(pos + e) - t.eval3(pos - e)
        assert np.array_equal(nrm[:, dim].view(np.uint32), want.view(np.uint32))
    # raw differences, not normalised (gleval.go:51-52): |n| ~ step for a unit-gradient field
    assert 0.5 * step < np.median(np.linalg.norm(nrm, axis=1)) < 1.5 * step
    with pytest.raises(gsdf_b200.GsdfError, match="invalid step"):
        gleval.NormalsCentralDiff(sdf, pos, nrm, 0.0)
    with pytest.raises(gsdf_b200.GsdfError, match="must match"):
        gleval.NormalsCentralDiff(sdf, pos, nrm[:10], step)
