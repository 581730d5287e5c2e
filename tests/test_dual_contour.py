"""Dual contouring (glrender/dual_contour.go, dual_contour_vertexplacement.go, gleval.NormalsCentralDiff).

CPU tests restate the reference's own property tests (glrender/dual_contour_test.go:140-497) on the oracle; GPU tests
require the CUDA renderer to reproduce the oracle's triangles bit for bit (same cube order, same float32 / float64
operation sequences) and re-check the same properties on its output."""
import numpy as np
import pytest

import gsdf_b200
from gsdf_b200 import gsdf, gleval, glrender


def surface_stats(tree, tris):
    v = np.unique(tris.reshape(-1, 3), axis=0)
    d = np.abs(tree.eval3(v))
    return float(d.max()), float(d.mean()), len(v)


def is_watertight(tris):
    """Every undirected edge of the triangle soup is used exactly twice, once in each direction."""
    v, inv = np.unique(tris.reshape(-1, 3), axis=0, return_inverse=True)
    f = inv.reshape(-1, 3)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    key = e[:, 0].astype(np.int64) * len(v) + e[:, 1]
    rkey = e[:, 1].astype(np.int64) * len(v) + e[:, 0]
    return len(np.unique(key)) == len(key) and np.array_equal(np.sort(key), np.sort(rkey))


SHAPES = {
    "sphere": lambda b: b.NewSphere(1.0),
    "box": lambda b: b.NewBox(2.0, 2.0, 2.0, 0.0),
    "snowman": lambda b: b.Union(b.NewSphere(0.4), b.Translate(b.NewSphere(0.3), 0, 0, 0.45), b.Translate(b.NewSphere(0.2), 0, 0, 0.8)),
}


# ---------------------------------------------------------------------------------------------- oracle (CPU)
def _qef_rows(points, normals, origin, lam):
    """The system DualContourLeastSquares hands to leastSquaresMGS64 (dual_contour_vertexplacement.go:107-133): one row
    n_i . x = n_i . (p_i - origin) per plane plus three regularisation rows sqrt(lam) * (x - bias)."""
    p = np.asarray(points, np.float32) - np.asarray(origin, np.float32)
    n = np.asarray(normals, np.float32)
    n = n / np.linalg.norm(n, axis=1, keepdims=True).astype(np.float32)
    sl = np.float32(np.sqrt(lam))
    bias = p.mean(axis=0, dtype=np.float32)
    A = np.concatenate([n, sl * np.eye(3, dtype=np.float32)])
    b = np.concatenate([(n * p).sum(axis=1, dtype=np.float32), sl * bias])
    return A.astype(np.float32), b.astype(np.float32)


def test_qef_solver_known_answers(oracle):
    """TestQEFSolver / TestQEFSolverDiagonalPlanes (dual_contour_test.go:20-137): three orthogonal planes meet at
    (0.5, 0.5, 0.5) (tolerance 1e-4, lambda 1e-6); three diagonal planes through (1, 1, 1) seen from the cube origin
    (0.5, 0.5, 0.5) (tolerance 1e-3, lambda 1e-8). The reference test solves the normal equations by a 3x3 inverse;
    the renderer's own solver is leastSquaresMGS64, restated in the oracle -- both must find the same points."""
    A, b = _qef_rows([[0.5, 0, 0], [0, 0.5, 0], [0, 0, 0.5]], [[1, 0, 0], [0, 1, 0], [0, 0, 1]], [0, 0, 0], 1e-6)
    x = oracle.lsq_mgs64(A, b)
    assert np.abs(x - np.float32(0.5)).max() < 1e-4
    org = np.array([0.5, 0.5, 0.5], np.float32)
    A, b = _qef_rows([[1, 1, 1]] * 3, [[1, 1, 0], [0, 1, 1], [1, 0, 1]], org, 1e-8)
    x = oracle.lsq_mgs64(A, b) + org
    assert np.abs(x - np.float32(1.0)).max() < 1e-3
    # fewer than three rows: the zero vector (dual_contour_vertexplacement.go:151-153); rank-deficient columns are zeroed
    assert np.array_equal(oracle.lsq_mgs64(A[:2], b[:2]), np.zeros(3, np.float32))
    x = oracle.lsq_mgs64([[1, 0, 0], [1, 0, 0], [0, 1, 0]], [2, 2, 3])
    assert np.allclose(x, [2, 3, 0], atol=1e-6)

def test_sphere_vertices_on_surface(oracle, bld):
    """TestDualContourSphereVerticesOnSurface (dual_contour_test.go:140-221): r=1, res=r/8."""
    s = bld.NewSphere(1.0)
    t = oracle.Tree.from_shader(s)
    res = np.float32(1.0 / 8)
    tris, st = oracle.dual_contour(t, *s.Bounds(), res, oracle.DC_LSQ)
    assert len(tris) > 0 and st["levels"] == 5  # bounds 2 wide / res -> log2(16)=4 -> 5 levels, 16^3 cubes
    mx, avg, _ = surface_stats(t, tris)
    tol = float(res) * 1.5
    assert mx <= tol and avg <= tol / 4
    # Not watertight at these parameters, faithfully: Reset translates the bounds by -res/2 (dual_contour.go:31-32) and
    # 2/res is an exact power of two, so the 16^3 cubes end at +0.9375 and the +x/+y/+z caps of the sphere get no quads.
    assert not is_watertight(tris)


def test_box_vertices_on_surface(oracle, bld):
    """TestDualContourBoxVerticesOnSurface (dual_contour_test.go:224-295): 2x2x2 box, res=size/8."""
    s = bld.NewBox(2.0, 2.0, 2.0, 0.0)
    t = oracle.Tree.from_shader(s)
    tris, _ = oracle.dual_contour(t, *s.Bounds(), np.float32(2.0 / 8), oracle.DC_LSQ)
    assert len(tris) > 0
    mx, _, _ = surface_stats(t, tris)
    assert mx <= 2.0 / 8 * 1.5


def test_least_squares_not_worse_than_naive(oracle, bld):
    """TestDualContourCompareWithNaive (dual_contour_test.go:300-351): r=1, res=r/6."""
    s = bld.NewSphere(1.0)
    t = oracle.Tree.from_shader(s)
    res = np.float32(1.0 / 6)
    lsq, _ = oracle.dual_contour(t, *s.Bounds(), res, oracle.DC_LSQ)
    naive, _ = oracle.dual_contour(t, *s.Bounds(), res, oracle.DC_NAIVE)
    assert len(lsq) == len(naive) > 0  # placement moves vertices, not topology
    assert is_watertight(lsq) and is_watertight(naive)  # 16 cubes of 1/6 cover 2.67 > 2: the surface closes
    assert surface_stats(t, lsq)[1] <= surface_stats(t, naive)[1] + 0.01


def test_no_suspicious_vertices(oracle, bld):
    """TestDualContourMissingNeighbors (dual_contour_test.go:429-497): nothing deep inside the sphere, nothing far outside."""
    s = bld.NewSphere(1.0)
    res = np.float32(1.0 / 8)
    tris, _ = oracle.dual_contour(oracle.Tree.from_shader(s), *s.Bounds(), res, oracle.DC_LSQ)
    v = tris.reshape(-1, 3)
    mn, mx = s.Bounds()
    assert (np.linalg.norm(v, axis=1) >= 0.5).all()
    assert (v >= mn - res).all() and (v <= mx + res).all()


def test_chiseled_uses_real_gradients(oracle, bld):
    """The default step 2e-8 vanishes in float32 for |coordinate| >= 0.25 (normals mostly zero -> the regularisation
    rows place the vertex at the mean crossing); Chiseled (step 1e-4) sees gradients and sharpens the box edges."""
    s = bld.NewBox(1.0, 2.0, 0.5, 0.0)  # TestBenchmarkBox (dual_contour_test.go:534-561)
    t = oracle.Tree.from_shader(s)
    res = np.float32(3.0 / 64)
    plain, _ = oracle.dual_contour(t, *s.Bounds(), res, oracle.DC_LSQ)
    chis, _ = oracle.dual_contour(t, *s.Bounds(), res, oracle.DC_LSQ_CHISELED)
    assert len(plain) == len(chis) > 0
    assert surface_stats(t, chis)[1] < surface_stats(t, plain)[1]
    assert is_watertight(chis)


def test_resolution_errors(oracle, bld):
    s = bld.NewSphere(1.0)
    mn, mx = s.Bounds()
    a = (np.ctypeslib.ctypes.c_float * 3)(*mn)
    b = (np.ctypeslib.ctypes.c_float * 3)(*mx)
    assert oracle.lib().go_dc_levels(a, b, 4.0, None) < 0   # "resolution not fine enough for marching cubes"
    assert oracle.lib().go_dc_levels(a, b, 0.0, None) < 0   # "invalid renderer cube resolution"
    assert oracle.lib().go_dc_levels(a, b, 0.125, None) == 5


def test_dual_render_bolt(oracle, bld):
    """TestDualRender (glrender_test.go:22-53): the rotated M3 bolt at res 0.5 through DualContourLeastSquares gives a
    non-empty mesh that survives the STL writer; quads come as triangle pairs."""
    s = gsdf.scene(bld, "bolt")
    t = oracle.Tree.from_shader(s)
    tris, st = oracle.dual_contour(t, *s.Bounds(), np.float32(0.5), oracle.DC_LSQ)
    assert len(tris) > 0 and len(tris) % 2 == 0 and np.isfinite(tris).all()
    mn, mx = s.Bounds()
    assert (tris.reshape(-1, 3) >= mn - 1.0).all() and (tris.reshape(-1, 3) <= mx + 1.0).all()
    stl = oracle.stl_write(tris)
    assert len(stl) == 84 + 50 * len(tris)


# ---------------------------------------------------------------------------------------------- CUDA (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("shape,res", [("sphere", 1.0 / 8), ("sphere", 1.0 / 6), ("box", 2.0 / 8), ("snowman", 3.0 / 64)])
@pytest.mark.parametrize("placer", ["naive", "lsq", "chiseled"])
def test_gpu_dual_contour_bit_identical(oracle, bld, shape, res, placer):
    s = SHAPES[shape](bld)
    sdf = gleval.NewCUDASDF3(s)
    vp = {"naive": glrender.DualContourNaive(), "lsq": glrender.DualContourLeastSquares(), "chiseled": glrender.DualContourLeastSquares(Chiseled=True)}[placer]
    okind = {"naive": oracle.DC_NAIVE, "lsq": oracle.DC_LSQ, "chiseled": oracle.DC_LSQ_CHISELED}[placer]
    dcr = glrender.DualContourRenderer()
    dcr.Reset(sdf, np.float32(res), vp)
    tris = dcr.RenderAll(None)
    want, st = oracle.dual_contour(oracle.Tree.from_shader(s), *s.Bounds(), np.float32(res), okind)
    got = dcr.Stats()
    assert (got["levels"], got["cubes"], got["with_neighbors"], got["evals"]) == (st["levels"], st["cubes"], st["with_neighbors"], st["evals"])
    assert len(tris) == len(want)
    assert np.array_equal(tris.view(np.uint32), want.view(np.uint32)), int((tris.view(np.uint32) != want.view(np.uint32)).any(axis=(1, 2)).sum())
    dcr.Rerun()
    assert np.array_equal(dcr.RenderAll(None).view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_gpu_dual_contour_flange(oracle, bld):
    """A benchmark scene (screw threads, smooth union): 64^3 cubes, bit-identical, watertight where closed."""
    s = gsdf.scene(bld, "npt-flange")
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(s.Diagonal() / np.float32(60))
    dcr = glrender.DualContourRenderer()
    dcr.Reset(sdf, res, glrender.DualContourLeastSquares(Chiseled=True))
    tris = dcr.RenderAll(None)
    want, st = oracle.dual_contour(oracle.Tree.from_shader(s), *s.Bounds(), res, oracle.DC_LSQ_CHISELED)
    assert dcr.Stats()["cubes"] == st["cubes"] and len(tris) == len(want) > 5000
    assert np.array_equal(tris.view(np.uint32), want.view(np.uint32))
    # RenderAll appends to dst (dual_contour.go:213-218)
    both = dcr.RenderAll(tris[:7])
    assert len(both) == len(tris) + 7 and np.array_equal(both[7:], tris)
    # the mesh goes through WriteBinarySTL like any other (dual_contour_test.go:529-531)
    import io
    buf = io.BytesIO()
    glrender.WriteBinarySTL(buf, tris)
    assert buf.getvalue() == oracle.stl_write(tris)


@pytest.mark.gpu
def test_gpu_dual_contour_properties_at_fine_resolution(bld):
    """256^3 cubes (16.7 M origin evaluations): the reference's property tests on the CUDA output alone."""
    s = bld.NewSphere(1.0)
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(2.0 / 200)
    dcr = glrender.DualContourRenderer()
    dcr.Reset(sdf, res, glrender.DualContourLeastSquares(Chiseled=True))
    tris = dcr.RenderAll(None)
    st = dcr.Stats()
    assert st["levels"] == 9 and st["triangles"] == len(tris) > 100000
    v = np.unique(tris.reshape(-1, 3), axis=0)
    d = np.abs(np.linalg.norm(v.astype(np.float64), axis=1) - 1.0)
    assert d.max() <= 1.5 * res and d.mean() <= 1.5 * res / 4
    assert is_watertight(tris)


@pytest.mark.gpu
def test_gpu_dual_contour_errors(bld):
    sdf = gleval.NewCUDASDF3(bld.NewSphere(1.0))
    dcr = glrender.DualContourRenderer()
    with pytest.raises(gsdf_b200.GsdfError, match="nil DualContourer"):
        dcr.Reset(sdf, 0.1, None)  # dual_contour.go:28-30
    with pytest.raises(gsdf_b200.GsdfError, match="resolution not fine enough"):
        dcr.Reset(sdf, 4.0, glrender.DualContourNaive())
    with pytest.raises(gsdf_b200.GsdfError, match="invalid renderer cube resolution"):
        dcr.Reset(sdf, 0.0, glrender.DualContourNaive())
    with pytest.raises(gsdf_b200.GsdfError, match="limit is 11 levels"):
        dcr.Reset(sdf, 1e-4, glrender.DualContourNaive())
    sdf2 = gleval.NewCUDASDF2(bld.NewCircle(1.0))
    with pytest.raises(gsdf_b200.GsdfError):
        dcr.Reset(sdf2, 0.1, glrender.DualContourNaive())


@pytest.mark.gpu
@pytest.mark.parametrize("shape,res,placer", [("sphere", 1.0 / 6, "lsq"), ("snowman", 3.0 / 64, "chiseled"), ("flange", None, "chiseled"), ("box", 2.0 / 8, "naive")])
def test_gpu_dual_contour_octant_parts_concatenate_to_the_whole(bld, shape, res, placer):
    """Multi-GPU layout: part r of nparts owns a run of top-level octants (a contiguous range of the BFS cube order) and
    recomputes a one-cube border on each side instead of exchanging halos; the parts' meshes concatenated in part order are
    bit-identical to the single-renderer mesh (same property as test_z_slabs_concatenate_to_the_whole for marching cubes)."""
    s = gsdf.scene(bld, "npt-flange") if shape == "flange" else SHAPES[shape](bld)
    if res is None:
        res = s.Diagonal() / np.float32(90)
    sdf = gleval.NewCUDASDF3(s)
    vp = {"naive": glrender.DualContourNaive(), "lsq": glrender.DualContourLeastSquares(), "chiseled": glrender.DualContourLeastSquares(Chiseled=True)}[placer]
    whole = glrender.DualContourRenderer()
    whole.Reset(sdf, np.float32(res), vp)
    want = whole.RenderAll(None)
    assert len(want) > 0
    for nparts in (2, 4, 8):
        parts, nb = [], 0
        for r in range(nparts):
            d = glrender.DualContourRenderer()
            d.Reset(sdf, np.float32(res), vp, part=r, nparts=nparts)
            parts.append(d.RenderAll(None))
            nb += d.Stats()["with_neighbors"]
            if nparts == 8:
                assert d.Stats()["evals"] < whole.Stats()["evals"]  # an octant plus its border, not the whole cube
            d.Close()
        got = np.concatenate(parts)
        assert len(got) == len(want), (nparts, [len(p) for p in parts])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), nparts
        assert nb == whole.Stats()["with_neighbors"]
    with pytest.raises(gsdf_b200.GsdfError, match="nparts must be"):
        glrender.DualContourRenderer().Reset(sdf, np.float32(res), vp, part=0, nparts=3)
